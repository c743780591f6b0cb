// Error reporting shared by the translation units of the library (not exported: -fvisibility=hidden).
#pragma once
// Records the message ntr_last_error() returns on this thread and hands `code` back (defined in capi.cu).
int ntr_fail(int code, const char *fmt, ...);
