/* device.cu -- see device.hpp */
#include "device.hpp"
#include <stdio.h>

namespace sqb {

void throwOnCudaError(cudaError_t st, const char *file, int line, const char *expr) {
    if (st == cudaSuccess) return;
    sq::throwErrorAt(file, line, "CUDA error %d (%s) in %s", (int)st, cudaGetErrorString(st), expr);
}

B200Device::B200Device() : launchCount(0), devNo_(-1), stream_(NULL), ownStream_(NULL), numSMs_(0), smemOptin_(0) {}
B200Device::~B200Device() {
    try { finalize(); } catch (...) {}
}

void B200Device::initialize(int devNo) {
    sqb_throwErrorIf(devNo_ >= 0, "Device already initialized.");
    int count = 0;
    cudaError_t st = cudaGetDeviceCount(&count);
    sqb_throwErrorIf(st != cudaSuccess || count == 0,
                     "no CUDA device available (%s); sqaod_b200 has no CPU fallback.", cudaGetErrorString(st));
    if (devNo < 0) devNo = 0;
    sqb_throwErrorIf(devNo >= count, "device %d not found (%d devices).", devNo, count);
    CUDA_CHECK(cudaSetDevice(devNo));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, devNo));
    sqb_throwErrorIf(prop.major < 10, "sqaod_b200 is built for sm_100a (B200); device %d is sm_%d%d.", devNo, prop.major, prop.minor);
    numSMs_ = prop.multiProcessorCount;
    smemOptin_ = prop.sharedMemPerBlockOptin;
    CUDA_CHECK(cudaStreamCreateWithFlags(&ownStream_, cudaStreamNonBlocking));
    stream_ = ownStream_;
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, devNo));
    uint64_t thr = UINT64_MAX; /* keep freed blocks in the pool: solvers re-prepare often */
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    devNo_ = devNo;
    sq::log("sqaod_b200: device %d %s, %d SMs, %zu KB smem/block", devNo, prop.name, numSMs_, smemOptin_ >> 10);
}

void B200Device::finalize() {
    if (devNo_ < 0) return;
    cudaSetDevice(devNo_);
    if (ownStream_) {
        cudaStreamSynchronize(ownStream_);
        cudaStreamDestroy(ownStream_);
    }
    ownStream_ = stream_ = NULL;
    devNo_ = -1;
}

void B200Device::makeCurrent() const {
    sqb_throwErrorIf(devNo_ < 0, "Device not initialized.");
    CUDA_CHECK(cudaSetDevice(devNo_));
}
void B200Device::setExternalStream(cudaStream_t s) {
    synchronize();
    stream_ = s ? s : ownStream_;
}
void B200Device::synchronize() const {
    makeCurrent();
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}
void *B200Device::alloc(size_t bytes) const {
    makeCurrent();
    void *p = NULL;
    CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 16, stream_));
    CUDA_CHECK(cudaMemsetAsync(p, 0, bytes ? bytes : 16, stream_));
    return p;
}
void B200Device::free(void *p) const {
    if (!p || devNo_ < 0) return;
    cudaSetDevice(devNo_);
    cudaFreeAsync(p, stream_);
}
void *B200Device::allocPinned(size_t bytes) const {
    makeCurrent();
    void *p = NULL;
    CUDA_CHECK(cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault));
    return p;
}
void B200Device::freePinned(void *p) const {
    if (p) cudaFreeHost(p);
}
void B200Device::h2d(void *dst, const void *src, size_t bytes) const {
    if (bytes) { makeCurrent(); CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_)); }
}
void B200Device::d2h(void *dst, const void *src, size_t bytes) const {
    if (bytes) { makeCurrent(); CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream_)); }
}
void B200Device::h2d2D(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height) const {
    if (width && height) { makeCurrent(); CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, stream_)); }
}
void B200Device::d2h2D(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height) const {
    if (width && height) { makeCurrent(); CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, stream_)); }
}

B200Device &asB200(sq::cuda::Device &dev) {
    B200Device *d = dynamic_cast<B200Device *>(&dev);
    sqb_throwErrorIf(d == NULL, "not a sqaod_b200 device.");
    sqb_throwErrorIf(!d->initialized(), "Device not initialized.");
    return *d;
}

} // namespace sqb

namespace sqaod { namespace cuda {
Device *newDevice(int devNo) {
    sqb::B200Device *d = new sqb::B200Device();
    if (devNo >= 0) {
        try { d->initialize(devNo); } catch (...) { delete d; throw; }
    }
    return d;
}
}} // namespace sqaod::cuda
