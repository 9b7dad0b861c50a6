"""Import alias: the package directory is ``fpv-drone-rl-agent_b200/`` (the name the
project brief fixes), which is not a valid Python identifier.  This shim makes
``import fpv_drone_rl_agent_b200`` resolve to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "fpv-drone-rl-agent_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
