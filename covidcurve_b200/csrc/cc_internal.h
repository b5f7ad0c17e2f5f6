// cc_internal.h - what the translation units of libcovidcurve_b200.so share besides the public header
#pragma once
#include <string>

// records the message cc_last_error() returns on this thread and hands back `code` (defined in covidcurve_b200.cu)
int cc_set_error(int code, const std::string& msg);
