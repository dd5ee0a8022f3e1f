// orbfe_host.h -- host-side helpers shared by the translation units of liborbfe.so.
#pragma once
#include <cstdarg>

// records a thread-local message for orbfe_last_error() and returns `code`
int orbfe_fail(int code, const char* fmt, ...);
