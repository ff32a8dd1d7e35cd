// lisa_b200/csrc/internal.h — helpers shared by the translation units of liblisa_rt.so that are not part of the C ABI.
#pragma once

// sets the calling thread's lisa_last_error() text (lisa_rt.cu owns the buffer)
extern "C" void lisa_internal_set_last_error(const char* msg);
