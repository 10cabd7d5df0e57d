// Version / error strings / launch counter of libpolcue.so.
#include "polcue_host.h"

extern "C" {

const char* polcue_version(void) { return "polcue 0.1.0 (sm_100a)"; }

const char* polcue_error_string(int code) {
    switch (code) {
        case POLCUE_OK: return "ok";
        case POLCUE_EINVAL: return "invalid argument (null pointer, non-positive or odd size, or misaligned buffer)";
        case POLCUE_ENOMEM: return "out of device or pinned memory";
        case POLCUE_ERANGE: return "refractive index whose zenith tables cannot be represented";
        case POLCUE_E2BIG: return "problem exceeds the work-item range of one launch; split the batch";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown polcue error";
    }
}

unsigned long long polcue_launch_count(void) { return polcue::g_launches.load(); }

}  // extern "C"
