// qx_internal.h -- shared between the translation units of libquadx_b200.so (not part of the public ABI)
#pragma once
// record the calling thread's last error text (read back with qx_last_error) and return `code`
int qx_fail(int code, const char* fmt, const char* detail);
