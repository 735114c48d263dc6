/* device.hpp -- device handle of the B200 back end (internal header).
 * Replaces the reference's Device/DeviceStream/DeviceMemoryStore/DeviceObjectAllocator stack
 * (sqaodc/cuda/Device.cpp:19-33, DeviceStream.cpp:32-68, DeviceMemoryStore.cpp) with: one CUDA stream per
 * Device, stream-ordered pool allocations (cudaMallocAsync), pinned staging buffers. */
#pragma once
#include <sqaod_b200/sqaod_api.hpp>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqb {

namespace sq = sqaod;

void throwOnCudaError(cudaError_t st, const char *file, int line, const char *expr);
#define CUDA_CHECK(expr) ::sqb::throwOnCudaError((expr), __FILE__, __LINE__, #expr)

class B200Device : public sq::cuda::Device {
public:
    B200Device();
    ~B200Device();
    int devNo() const { return devNo_; }
    void initialize(int devNo = 0);
    void finalize();

    bool initialized() const { return devNo_ >= 0; }
    void makeCurrent() const;
    /* the stream every kernel of this Device is launched on; selects the device first, so that solvers assigned to
     * different devices can be interleaved in one process (every launch site evaluates dev.stream()) */
    cudaStream_t stream() const { makeCurrent(); return stream_; }
    /* bench.py / torch interop: run on a caller-provided stream (e.g. torch's current stream). */
    void setExternalStream(cudaStream_t s);
    void synchronize() const;
    int numSMs() const { return numSMs_; }
    size_t smemPerBlockOptin() const { return smemOptin_; }

    void *alloc(size_t bytes) const;           /* stream-ordered, zero-filled */
    void free(void *p) const;
    void *allocPinned(size_t bytes) const;
    void freePinned(void *p) const;
    void h2d(void *dst, const void *src, size_t bytes) const;   /* async on stream(), src may be pageable */
    void d2h(void *dst, const void *src, size_t bytes) const;
    void h2d2D(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height) const;
    void d2h2D(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height) const;

    /* launches of this library's own kernels since the last reset (bench.py "gpu_launches") */
    mutable unsigned long long launchCount;

private:
    int devNo_;
    cudaStream_t stream_, ownStream_;
    int numSMs_;
    size_t smemOptin_;
};

B200Device &asB200(sq::cuda::Device &dev);

template <class T> struct DevBuf { /* RAII device array tied to a device */
    DevBuf() : dev(NULL), p(NULL), n(0) {}
    ~DevBuf() { release(); }
    void alloc(const B200Device *d, size_t count) { release(); dev = d; n = count; p = (T *)d->alloc(sizeof(T) * (count ? count : 1)); }
    void release() { if (p && dev) dev->free(p); p = NULL; n = 0; }
    const B200Device *dev;
    T *p;
    size_t n;
private:
    DevBuf(const DevBuf &);
    DevBuf &operator=(const DevBuf &);
};

} // namespace sqb
