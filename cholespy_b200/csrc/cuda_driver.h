// CUDA driver shim for the B200 solve path: dlopen(libcuda), resolve entry points, load the
// embedded sm_100a cubin, one primary context + module per device.
//
// Replaces the reference's shim (reference: src/cuda_driver.h:1-61, src/cuda_driver.cpp:58-149).
// The names the reference's unmodified bindings use (src/main.cpp:51-61,102) are kept with the same
// meaning so that file links against this one: init_cuda(), shutdown_cuda(), cu_context,
// scoped_set_context, cuda_check / cuda_check_impl, the cuMemcpyAsync function pointer and the
// CUdeviceptr = void* typedef.  No CUDA headers are needed to build the host side.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>

#define CUDA_SUCCESS 0
#define CUDA_ERROR_DEINITIALIZED 4

using CUcontext = struct CUctx_st *;
using CUmodule = struct CUmod_st *;
using CUstream = struct CUstream_st *;
using CUevent = struct CUevent_st *;
using CUfunction = struct CUfunc_st *;
using CUresult = int;
using CUdevice = int;
using CUdeviceptr = void *;

// cuMemcpy2DAsync descriptor (driver ABI layout of CUDA_MEMCPY2D; memory types: 1 = host, 2 = device)
struct CUDA_MEMCPY2D {
    size_t srcXInBytes, srcY;
    unsigned int srcMemoryType;
    const void *srcHost;
    CUdeviceptr srcDevice;
    void *srcArray;
    size_t srcPitch;
    size_t dstXInBytes, dstY;
    unsigned int dstMemoryType;
    void *dstHost;
    CUdeviceptr dstDevice;
    void *dstArray;
    size_t dstPitch;
    size_t WidthInBytes, Height;
};

// -- entry points (resolved by init_cuda) ---------------------------------------------------------
extern CUresult (*cuInit)(unsigned int);
extern CUresult (*cuDeviceGet)(CUdevice *, int);
extern CUresult (*cuDeviceGetCount)(int *);
extern CUresult (*cuDeviceGetAttribute)(int *, int, CUdevice);
extern CUresult (*cuDevicePrimaryCtxRetain)(CUcontext *, CUdevice);
extern CUresult (*cuDevicePrimaryCtxRelease)(CUdevice);
extern CUresult (*cuCtxPushCurrent)(CUcontext);
extern CUresult (*cuCtxPopCurrent)(CUcontext *);
extern CUresult (*cuCtxSynchronize)(void);
extern CUresult (*cuGetErrorName)(CUresult, const char **);
extern CUresult (*cuGetErrorString)(CUresult, const char **);
extern CUresult (*cuMemAlloc)(void **, size_t);
extern CUresult (*cuMemFree)(void *);
extern CUresult (*cuMemHostAlloc)(void **, size_t, unsigned int);
extern CUresult (*cuMemFreeHost)(void *);
extern CUresult (*cuMemcpyAsync)(void *, const void *, size_t, CUstream);
extern CUresult (*cuMemsetD8Async)(void *, unsigned char, size_t, CUstream);
extern CUresult (*cuMemsetD32Async)(void *, unsigned int, size_t, CUstream);
extern CUresult (*cuModuleLoadData)(CUmodule *, const void *);
extern CUresult (*cuModuleGetFunction)(CUfunction *, CUmodule, const char *);
extern CUresult (*cuModuleUnload)(CUmodule);
extern CUresult (*cuFuncSetAttribute)(CUfunction, int, int);
extern CUresult (*cuFuncGetAttribute)(int *, int, CUfunction);
extern CUresult (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t);
extern CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                  CUstream, void **, void **);
extern CUresult (*cuLaunchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                             unsigned, CUstream, void **);
extern CUresult (*cuStreamSynchronize)(CUstream);
extern CUresult (*cuStreamCreate)(CUstream *, unsigned int);
extern CUresult (*cuStreamDestroy)(CUstream);
extern CUresult (*cuStreamWaitEvent)(CUstream, CUevent, unsigned int);
extern CUresult (*cuEventCreate)(CUevent *, unsigned int);
extern CUresult (*cuEventDestroy)(CUevent);
extern CUresult (*cuEventRecord)(CUevent, CUstream);
extern CUresult (*cuMemcpy2DAsync)(const CUDA_MEMCPY2D *, CUstream);
extern CUresult (*cuPointerGetAttribute)(void *, int, CUdeviceptr);
extern CUresult (*cuDeviceCanAccessPeer)(int *, CUdevice, CUdevice);
extern CUresult (*cuCtxEnablePeerAccess)(CUcontext, unsigned int);
extern CUresult (*cuMemcpyPeerAsync)(CUdeviceptr, CUcontext, CUdeviceptr, CUcontext, size_t, CUstream);

#define cuda_check(err) cuda_check_impl(err, __FILE__, __LINE__)
// Prints CUDA API failures to stderr and records the first one for chb_last_error(); never throws
// (same policy as reference src/cuda_driver.cpp:48-56).
extern void cuda_check_impl(CUresult errval, const char *file, const int line);
extern CUresult cuda_first_error();  // sticky: first non-success code seen since the last clear
extern void cuda_clear_error();

// -- reference-compatible surface: device 0 -------------------------------------------------------
extern CUdevice cu_device;
extern CUcontext cu_context;
extern bool init_cuda();  // loads libcuda, retains device 0's primary context, loads the cubin
extern void shutdown_cuda();

struct scoped_set_context {
    scoped_set_context(CUcontext ctx) { cuda_check(cuCtxPushCurrent(ctx)); }
    ~scoped_set_context() { cuda_check(cuCtxPopCurrent(nullptr)); }
};

// -- per-device state (new: any device ordinal, for one-process-per-GPU launches) -----------------
namespace chb {

struct DeviceState {
    int ordinal = -1;
    CUdevice dev = 0;
    CUcontext ctx = nullptr;
    CUmodule module = nullptr;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
};

// Returns the initialised state for `ordinal` (retains its primary context and loads the module on
// first use) or nullptr if CUDA / the device / an sm_100 part is not available (error text is set).
DeviceState *device_state(int ordinal);
const char *driver_error_text();
// Kernel entry point by name from the device's module (cached); nullptr + error text if missing.
CUfunction kernel(DeviceState *ds, const char *name);
// Largest co-resident grid for `fn` at `block` threads / `smem` dynamic bytes (blocks per SM x SMs).
int max_resident_blocks(DeviceState *ds, CUfunction fn, int block, size_t smem);
// Device ordinal that owns a device pointer (CU_POINTER_ATTRIBUTE_DEVICE_ORDINAL), -1 if unknown.
int device_of_pointer(const void *dptr);

}  // namespace chb
