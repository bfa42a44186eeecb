// Version / error text / device query of the C ABI (include/lighthand_b200.h).
#include "lh_common.cuh"

extern "C" int lh_abi_version(void) { return LH_ABI_VERSION; }

extern "C" const char* lh_strerror(int code) {
    switch (code) {
        case LH_OK: return "ok";
        case LH_EINVAL: return "lighthand_b200: invalid argument (size / enum)";
        case LH_ENULL: return "lighthand_b200: required pointer is NULL";
        case LH_EALIGN: return "lighthand_b200: pointer or pitch alignment requirement not met";
        case LH_EUNSUPPORTED: return "lighthand_b200: geometry not supported by the sm_100a kernels";
        case LH_ENOMEM: return "lighthand_b200: out of memory / workspace too small";
        case LH_ENODEVICE: return "lighthand_b200: no CUDA device of compute capability 10.x";
        case LH_EJPEG: return "lighthand_b200: nvJPEG unavailable (libnvjpeg.so.12 / backend) or an undecodable bit-stream";
        case LH_ENCCL: return "lighthand_b200: NCCL unavailable (libnccl.so.2) or an NCCL call failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "lighthand_b200: unknown error";
}

extern "C" int lh_device_query(int device, int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return LH_ENODEVICE;
    if (device < 0 || device >= n) return LH_EINVAL;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem) *total_mem = (uint64_t)prop.totalGlobalMem;
    return LH_OK;
}
