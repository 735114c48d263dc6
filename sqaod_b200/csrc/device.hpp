/* device.hpp -- device handle of the B200 back end (internal header).
 * Replaces the reference's Device/DeviceStream/DeviceMemoryStore/DeviceObjectAllocator stack
 * (sqaodc/cuda/Device.cpp:19-33, DeviceStream.cpp:32-68, DeviceMemoryStore.cpp) with: one CUDA stream per
 * Device, stream-ordered pool allocations (cudaMallocAsync), pinned staging buffers. */
#pragma once
#include <sqaod_b200/sqaod_api.hpp>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

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

/* Asynchronous device -> host read-back of a small result (the energies): the copy lands in pinned host memory and is fenced by its
 * own event, so calculate_E() only ENQUEUES work and the launch stream is never synchronised; whoever reads the values (get_E)
 * waits for that event alone.  (The reference synchronises the whole device per query, CUDADenseGraphAnnealer.cu:185-192, 337-351.) */
template <class T> struct AsyncReadback {
    AsyncReadback() : dev(NULL), host(NULL), n(0), ev(NULL), pending(false) {}
    ~AsyncReadback() { release(); }
    void alloc(const B200Device *d, size_t count) {
        if (host && dev == d && n == count) { pending = false; return; } /* solvers re-prepare often: keep the pinned buffer */
        release();
        dev = d; n = count;
        host = (T *)d->allocPinned(sizeof(T) * (count ? count : 1));
        d->makeCurrent();
        CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    void release() {
        if (host && dev) dev->freePinned(host);
        if (ev) cudaEventDestroy(ev);
        host = NULL; ev = NULL; n = 0; pending = false;
    }
    void enqueue(const T *devPtr, size_t count) { /* after the kernels that produce devPtr, on the device's stream */
        dev->d2h(host, devPtr, sizeof(T) * count);
        CUDA_CHECK(cudaEventRecord(ev, dev->stream()));
        pending = true;
    }
    void wait(T *dst, size_t count) { /* no-op when nothing is in flight */
        if (!pending) return;
        CUDA_CHECK(cudaEventSynchronize(ev));
        memcpy(dst, host, sizeof(T) * count);
        pending = false;
    }
    const B200Device *dev;
    T *host;
    size_t n;
    cudaEvent_t ev;
    bool pending;
private:
    AsyncReadback(const AsyncReadback &);
    AsyncReadback &operator=(const AsyncReadback &);
};

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
