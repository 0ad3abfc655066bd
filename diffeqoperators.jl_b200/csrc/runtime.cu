// Runtime of libdeo_b200: error text, device selection, the library streams, device buffers.
#include "common.hpp"

namespace deo {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}
const std::string& last_error() { return t_error; }

int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();   // clear the sticky-less error state
    return DEO_ERR_CUDA;
}

Runtime& rt() {
    static Runtime r;
    return r;
}

int32_t ensure_init() {
    if (rt().ready) return DEO_OK;
    return deo_init(0);
}

}  // namespace deo

using namespace deo;

extern "C" {

int32_t deo_abi_version(void) { return DEO_ABI_VERSION; }

int32_t deo_device_count(int32_t* count) {
    DEO_REQUIRE(count != nullptr, "deo_device_count: null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    *count = n;
    return DEO_OK;
}

int32_t deo_init(int32_t device) {
    Runtime& r = rt();
    if (r.ready && r.device == device) return DEO_OK;
    int n = 0;
    DEO_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) { set_error("deo_init: no CUDA device (this library has no CPU fallback)"); return DEO_ERR_CUDA; }
    DEO_REQUIRE(device >= 0 && device < n, "deo_init: device %d out of range (have %d)", device, n);
    DEO_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DEO_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("deo_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return DEO_ERR_UNSUPPORTED;
    }
    if (r.ready) {   // switching device: drop old streams
        cudaStreamDestroy(r.stream);
        cudaStreamDestroy(r.comm_stream);
        cudaStreamDestroy(r.h2d_stream);
        cudaStreamDestroy(r.d2h_stream);
        r.ready = false;
    }
    DEO_CUDA(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
    int lo = 0, hi = 0;
    DEO_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    DEO_CUDA(cudaStreamCreateWithPriority(&r.comm_stream, cudaStreamNonBlocking, hi));
    DEO_CUDA(cudaStreamCreateWithFlags(&r.h2d_stream, cudaStreamNonBlocking));
    DEO_CUDA(cudaStreamCreateWithFlags(&r.d2h_stream, cudaStreamNonBlocking));
    r.device = device;
    r.sm_count = prop.multiProcessorCount;
    r.ready = true;
    return DEO_OK;
}

// A slab launch that gave up waiting for its neighbours' halo planes raises a mapped host word instead of trapping
// (the context stays usable); the next synchronising call reports it once.
static int32_t check_halo_timeout() {
    if (star_take_halo_timeout()) {
        set_error("halo exchange timed out: a slab launch waited longer than the bound (DEO_HALO_TIMEOUT_S) for a neighbour's halo planes; its result is invalid");
        return DEO_ERR_CUDA;
    }
    return DEO_OK;
}

int32_t deo_sync(void) {
    if (!rt().ready) return DEO_OK;
    DEO_CUDA(cudaStreamSynchronize(rt().stream));
    DEO_CUDA(cudaStreamSynchronize(rt().comm_stream));
    return check_halo_timeout();
}

int32_t deo_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return DEO_ERR_INVALID;
    const std::string& e = last_error();
    size_t n = e.size() < len - 1 ? e.size() : len - 1;
    memcpy(buf, e.data(), n);
    buf[n] = 0;
    return DEO_OK;
}

int32_t deo_launch_count(int64_t* count) {
    DEO_REQUIRE(count != nullptr, "deo_launch_count: null argument");
    *count = (int64_t)g_launches.load();
    return DEO_OK;
}

// ---- buffers -----------------------------------------------------------------------------------
int32_t deo_buffer_create(size_t bytes, deo_buffer** out) {
    DEO_REQUIRE(out != nullptr, "deo_buffer_create: null argument");
    *out = nullptr;
    int32_t rc = ensure_init();
    if (rc) return rc;
    deo_buffer* b = new (std::nothrow) deo_buffer();
    if (!b) { set_error("out of host memory"); return DEO_ERR_NOMEM; }
    cudaError_t e = cudaMalloc(&b->ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        delete b;
        cudaGetLastError();
        set_error("deo_buffer_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? DEO_ERR_NOMEM : DEO_ERR_CUDA;
    }
    b->bytes = bytes;
    b->owned = true;
    *out = b;
    return DEO_OK;
}

int32_t deo_buffer_wrap(void* devptr, size_t bytes, deo_buffer** out) {
    DEO_REQUIRE(out != nullptr && devptr != nullptr, "deo_buffer_wrap: null argument");
    int32_t rc = ensure_init();
    if (rc) return rc;
    deo_buffer* b = new (std::nothrow) deo_buffer();
    if (!b) { set_error("out of host memory"); return DEO_ERR_NOMEM; }
    b->ptr = devptr; b->bytes = bytes; b->owned = false;
    *out = b;
    return DEO_OK;
}

int32_t deo_buffer_free(deo_buffer* buf) {
    if (!buf) return DEO_OK;
    if (buf->owned && buf->ptr) {
        // nothing of ours may still be reading or writing the allocation: the compute stream, the halo pushes / NCCL
        // transfers of the communication stream and the chunk pipeline's copy streams
        if (rt().ready) {
            cudaStreamSynchronize(rt().stream);
            cudaStreamSynchronize(rt().comm_stream);
            cudaStreamSynchronize(rt().h2d_stream);
            cudaStreamSynchronize(rt().d2h_stream);
        }
        deo::dist_forget_buffer(buf->ptr);
        cudaFree(buf->ptr);
    }
    delete buf;
    return DEO_OK;
}

int32_t deo_buffer_size(const deo_buffer* buf, size_t* bytes) {
    DEO_REQUIRE(buf && bytes, "deo_buffer_size: null argument");
    *bytes = buf->bytes;
    return DEO_OK;
}

int32_t deo_buffer_upload(deo_buffer* dst, const void* host, size_t bytes) {
    DEO_REQUIRE(dst && host, "deo_buffer_upload: null argument");
    DEO_REQUIRE(bytes <= dst->bytes, "deo_buffer_upload: %zu bytes into a %zu-byte buffer", bytes, dst->bytes);
    DEO_CUDA(cudaMemcpyAsync(dst->ptr, host, bytes, cudaMemcpyHostToDevice, rt().stream));
    DEO_CUDA(cudaStreamSynchronize(rt().stream));   // the host array may be reused immediately
    return DEO_OK;
}

int32_t deo_buffer_download(void* host, const deo_buffer* src, size_t bytes) {
    DEO_REQUIRE(src && host, "deo_buffer_download: null argument");
    DEO_REQUIRE(bytes <= src->bytes, "deo_buffer_download: %zu bytes from a %zu-byte buffer", bytes, src->bytes);
    DEO_CUDA(cudaMemcpyAsync(host, src->ptr, bytes, cudaMemcpyDeviceToHost, rt().stream));
    DEO_CUDA(cudaStreamSynchronize(rt().stream));
    return check_halo_timeout();
}

int32_t deo_buffer_devptr(const deo_buffer* buf, void** devptr) {
    DEO_REQUIRE(buf && devptr, "deo_buffer_devptr: null argument");
    *devptr = buf->ptr;
    return DEO_OK;
}

int32_t deo_host_alloc(size_t bytes, void** host) {
    DEO_REQUIRE(host != nullptr, "deo_host_alloc: null argument");
    int32_t rc = ensure_init();
    if (rc) return rc;
    DEO_CUDA(cudaMallocHost(host, bytes ? bytes : 1));
    return DEO_OK;
}

int32_t deo_host_free(void* host) {
    if (host) DEO_CUDA(cudaFreeHost(host));
    return DEO_OK;
}

}  // extern "C"
