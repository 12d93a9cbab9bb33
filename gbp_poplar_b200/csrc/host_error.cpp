// Error text shared by the host setup code and the CUDA C ABI
// (gbp_cuda_last_error, include/gbp_cuda.h).
#include <string>

#include "../../include/gbp_cuda.h"

static thread_local std::string g_gbp_error;

void gbp_set_error(const std::string& s) { g_gbp_error = s; }

extern "C" const char* gbp_cuda_last_error(void) { return g_gbp_error.c_str(); }
