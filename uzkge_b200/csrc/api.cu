// api.cu -- the C ABI of include/uzkge_cuda.h: process-wide state, host<->device staging, error mapping.
//
// The Rust side binds exactly these symbols from the bodies of
// KZGCommitmentSchemeBN254::commit (/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:278-293) and
// FpPolynomial::{fft,ifft,coset_fft,coset_ifft}_with_domain (/root/reference/uzkge/src/poly_commit/field_polynomial.rs:583-607).
// There is no CPU path in this library: without a Blackwell device every call fails.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "devmem.cuh"
#include "internal.h"

namespace uz {
std::atomic<uint64_t> g_launches{0};
Profiler g_prof;
}

using namespace uz;

namespace uz {
extern int g_quotient_min_blocks;  // quotient.cu
}

namespace {

thread_local std::string t_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
};

struct State {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::unique_ptr<NttEngine> ntt;
    std::unique_ptr<MsmEngine> msm;
    std::unique_ptr<PolyEngine> poly;
    std::map<uint64_t, MsmSrs> srs;
    uint64_t next_handle = 1;
    DevBuf data, scratch, scalars, small;
    cudaStream_t copy_stream = nullptr, out_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    uint32_t ntt_log_tile = 10, ntt_max_log_r = 10, ntt_two_pass_max = 18;  // measured best on B200 (scripts/gpu_ntt_cfg.py)
};
State g;

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) {
        snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        cudaGetLastError();  // clear the sticky-free error state
    } else {
        snprintf(buf, sizeof buf, "%s", what);
    }
    t_error = buf;
    return code;
}
int fail_cuda(const char* what, cudaError_t e) { return fail(cuda_err_code(e), what, e); }

int engine_fail(int rc, const char* what) {
    if (rc == UZKGE_OK) return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(rc, what, e);
    return fail(rc, what);
}

// caller holds g.mu
int ensure_init(int device) {
    if (g.ready) {
        cudaError_t e = cudaSetDevice(g.device);
        if (e != cudaSuccess) return fail_cuda("cudaSetDevice", e);
        return UZKGE_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(UZKGE_ERR_NO_DEVICE, "no CUDA device: uzkge-b200 has no CPU path");
    }
    if (device >= 0) {
        if (device >= count) return fail(UZKGE_ERR_NO_DEVICE, "device index out of range");
        e = cudaSetDevice(device);
        if (e != cudaSuccess) return fail_cuda("cudaSetDevice", e);
    }
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail_cuda("cudaGetDevice", e);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail_cuda("cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        char buf[256];
        snprintf(buf, sizeof buf, "device %d (%.64s, sm_%d%d) is not a Blackwell sm_100 part: kernels are built for sm_100a only",
                 dev, prop.name, prop.major, prop.minor);
        return fail(UZKGE_ERR_NO_DEVICE, buf);
    }
    e = cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail_cuda("cudaStreamCreate", e);
    g.device = dev;
    g.sm_count = prop.multiProcessorCount;
    g.ntt.reset(new NttEngine(g.sm_count));
    g.ntt->configure(g.ntt_log_tile, g.ntt_max_log_r, g.ntt_two_pass_max);
    g.msm.reset(new MsmEngine(g.sm_count));
    g.poly.reset(new PolyEngine());
    g.ready = true;
    return UZKGE_OK;
}

#define API_ENTER(dev)                           \
    std::lock_guard<std::mutex> lock__(g.mu);    \
    do {                                         \
        int rc__ = ensure_init(dev);             \
        if (rc__ != UZKGE_OK) return rc__;       \
    } while (0)

#define CUDA_OR_FAIL(expr, what)                         \
    do {                                                 \
        cudaError_t e__ = (expr);                        \
        if (e__ != cudaSuccess) return fail_cuda(what, e__); \
    } while (0)

// ---- K1 test / roof kernels
template <class P>
__global__ void field_mul_kernel(const fe* a, const fe* b, fe* o, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fe(o + i, fe_mul<P>(ld_fe(a + i), ld_fe(b + i)));
}
// 4 independent dependent-chains per thread: x_k <- x_k * y_k, `iters` times
template <class P>
__global__ void __launch_bounds__(256) field_mul_bench_kernel(fe* sink, uint32_t iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fe x[4], y;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = fe_one<P>();
        x[k].l[0] += t * 4 + k;
    }
    y = fe_one<P>();
    y.l[1] ^= t;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] = fe_mul<P>(x[k], y);
    }
    fe r = fe_add<P>(fe_add<P>(x[0], x[1]), fe_add<P>(x[2], x[3]));
    if (r.l[7] == 0xffffffffu) st_fe(sink + t, r);  // never true for reduced values: keeps the chain alive
}

}  // namespace

extern "C" {

UZKGE_API int32_t uzkge_cuda_init(int32_t device) {
    API_ENTER(device);
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

UZKGE_API const char* uzkge_cuda_last_error(void) { return t_error.c_str(); }
UZKGE_API const char* uzkge_cuda_version(void) { return "uzkge-b200 0.1.0 sm_100a"; }

UZKGE_API int32_t uzkge_cuda_srs_upload(const uint64_t* affine_xy, size_t n, uint32_t window_bits, uint64_t* handle) {
    if (!affine_xy || !handle) return fail(UZKGE_ERR_ARG, "srs_upload: null pointer");
    API_ENTER(-1);
    MsmSrs s;
    int rc = g.msm->upload(affine_xy, n, window_bits, &s, g.stream);
    if (rc != UZKGE_OK) {
        g.msm->release(&s);
        return engine_fail(rc, "srs_upload");
    }
    const uint64_t h = g.next_handle++;
    g.srs[h] = s;
    *handle = h;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_srs_free(uint64_t handle) {
    API_ENTER(-1);
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "srs_free: unknown handle");
    cudaStreamSynchronize(g.stream);
    g.msm->release(&it->second);
    g.srs.erase(it);
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_srs_info(uint64_t handle, uzkge_srs_info* info) {
    if (!info) return fail(UZKGE_ERR_ARG, "srs_info: null pointer");
    API_ENTER(-1);
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "srs_info: unknown handle");
    info->window_bits = it->second.c;
    info->windows = it->second.windows;
    info->n = it->second.n;
    info->device_bytes = it->second.bytes;
    info->batch_slots = it->second.slots;
    info->reserved = 0;
    info->precompute_ms = it->second.precompute_ms;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_msm_g1_batch(uint64_t handle, const uint64_t* const* scalars, const size_t* n, size_t k, uint64_t* out_jac) {
    if (k && (!scalars || !n || !out_jac)) return fail(UZKGE_ERR_ARG, "msm_g1_batch: null pointer");
    API_ENTER(-1);
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1: unknown handle");
    if (k == 0) return UZKGE_OK;
    MsmSrs& s = it->second;
    size_t total = 0;
    for (size_t j = 0; j < k; j++) {
        if (n[j] > s.n) return fail(UZKGE_ERR_SIZE, "msm_g1: more scalars than SRS points");
        if (n[j] && !scalars[j]) return fail(UZKGE_ERR_ARG, "msm_g1: null scalar vector");
        total += n[j];
    }
    CUDA_OR_FAIL(g.scalars.reserve(total * sizeof(fe) + 32), "msm_g1: scalar buffer");
    CUDA_OR_FAIL(g.small.reserve(k * sizeof(jacobian) + 4096), "msm_g1: output buffer");
    fe* d_s = (fe*)g.scalars.p;
    jacobian* d_o = (jacobian*)g.small.p;
    // H2D copies on a copy stream, one event per vector: the engine's pipeline starts group g as soon as its scalars have
    // landed, so the copies of later groups run under the kernels of earlier ones
    if (!g.copy_stream) CUDA_OR_FAIL(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking), "msm_g1_batch: stream");
    while (g.copy_events.size() < k) {
        cudaEvent_t ev;
        CUDA_OR_FAIL(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "msm_g1_batch: event");
        g.copy_events.push_back(ev);
    }
    std::vector<const fe*> ptrs(k);
    size_t off = 0;
    for (size_t j = 0; j < k; j++) {
        if (n[j]) CUDA_OR_FAIL(cudaMemcpyAsync(d_s + off, scalars[j], n[j] * sizeof(fe), cudaMemcpyHostToDevice, g.copy_stream), "msm_g1: H2D");
        CUDA_OR_FAIL(cudaEventRecord(g.copy_events[j], g.copy_stream), "msm_g1_batch: event");
        ptrs[j] = d_s + off;
        off += n[j];
    }
    int rc = g.msm->run_pipelined(&s, 0, ptrs.data(), n, k, d_o, g.stream, g.copy_events.data());
    if (rc != UZKGE_OK) {
        cudaStreamSynchronize(g.copy_stream);
        cudaStreamSynchronize(g.stream);
        return engine_fail(rc, "msm_g1_batch: launch");
    }
    CUDA_OR_FAIL(cudaMemcpyAsync(out_jac, d_o, k * sizeof(jacobian), cudaMemcpyDeviceToHost, g.stream), "msm_g1: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "msm_g1: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_msm_g1(uint64_t handle, size_t base_offset, const uint64_t* scalars, size_t n, uint64_t out_jac[12]) {
    if (!out_jac || (n && !scalars)) return fail(UZKGE_ERR_ARG, "msm_g1: null pointer");
    API_ENTER(-1);
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1: unknown handle");
    MsmSrs& s = it->second;
    if (base_offset > s.n || n > s.n - base_offset) return fail(UZKGE_ERR_SIZE, "msm_g1: range outside the SRS");
    CUDA_OR_FAIL(g.scalars.reserve(n * sizeof(fe) + 32), "msm_g1: scalar buffer");
    CUDA_OR_FAIL(g.small.reserve(4096), "msm_g1: output buffer");
    if (n) CUDA_OR_FAIL(cudaMemcpyAsync(g.scalars.p, scalars, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "msm_g1: H2D");
    int rc = g.msm->run(&s, base_offset, (const fe*)g.scalars.p, n, (jacobian*)g.small.p, g.stream);
    if (rc != UZKGE_OK) {
        cudaStreamSynchronize(g.stream);
        return engine_fail(rc, "msm_g1: launch");
    }
    CUDA_OR_FAIL(cudaMemcpyAsync(out_jac, g.small.p, sizeof(jacobian), cudaMemcpyDeviceToHost, g.stream), "msm_g1: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "msm_g1: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_msm_g1_device(uint64_t handle, size_t base_offset, const void* d_scalars, size_t n, void* d_out_jac, void* stream) {
    if (!d_out_jac || (n && !d_scalars)) return fail(UZKGE_ERR_ARG, "msm_g1_device: null pointer");
    API_ENTER(-1);
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1_device: unknown handle");
    int rc = g.msm->run(&it->second, base_offset, (const fe*)d_scalars, n, (jacobian*)d_out_jac, (cudaStream_t)stream);
    return engine_fail(rc, "msm_g1_device");
}

UZKGE_API int32_t uzkge_cuda_msm_g1_batch_device(uint64_t handle, size_t base_offset, const void* const* d_scalars, const size_t* n,
                                                 size_t k, void* d_out_jac, void* stream) {
    if (k && (!d_scalars || !n || !d_out_jac)) return fail(UZKGE_ERR_ARG, "msm_g1_batch_device: null pointer");
    API_ENTER(-1);
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1_batch_device: unknown handle");
    int rc = g.msm->run_pipelined(&it->second, base_offset, (const fe* const*)d_scalars, n, k, (jacobian*)d_out_jac, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "msm_g1_batch_device: range outside the SRS");
    return engine_fail(rc, "msm_g1_batch_device");
}

UZKGE_API int32_t uzkge_cuda_ntt_fr(uint64_t* inout, size_t len_in, size_t domain_size, int32_t inverse, const uint64_t* coset_shift) {
    if (!inout) return fail(UZKGE_ERR_ARG, "ntt_fr: null pointer");
    if (inverse != 0 && inverse != 1) return fail(UZKGE_ERR_ARG, "ntt_fr: inverse must be 0 or 1");
    if (len_in > domain_size) return fail(UZKGE_ERR_SIZE, "ntt_fr: input longer than the domain");
    API_ENTER(-1);
    bool ok = false;
    ntt_root_of_unity(domain_size, &ok);
    if (!ok || domain_size % 9 == 0) return fail(UZKGE_ERR_SIZE, "ntt_fr: domain size must be 2^k or 3 * 2^k, k <= 28");
    CUDA_OR_FAIL(g.data.reserve(domain_size * sizeof(fe)), "ntt_fr: data buffer");
    CUDA_OR_FAIL(g.scratch.reserve(domain_size * sizeof(fe)), "ntt_fr: scratch buffer");
    if (len_in) CUDA_OR_FAIL(cudaMemcpyAsync(g.data.p, inout, len_in * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "ntt_fr: H2D");
    fe shift;
    if (coset_shift) memcpy(&shift, coset_shift, sizeof(fe));
    int rc = g.ntt->run((const fe*)g.data.p, (fe*)g.data.p, (fe*)g.scratch.p, len_in, domain_size, inverse != 0,
                        coset_shift ? &shift : nullptr, g.stream);
    if (rc != UZKGE_OK) {
        cudaStreamSynchronize(g.stream);
        return engine_fail(rc, "ntt_fr: launch");
    }
    CUDA_OR_FAIL(cudaMemcpyAsync(inout, g.data.p, domain_size * sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "ntt_fr: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "ntt_fr: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_ntt_fr_batch(uint64_t* const* inouts, const size_t* len_in, size_t k, size_t domain_size, int32_t inverse,
                                          const uint64_t* coset_shift) {
    if (k && (!inouts || !len_in)) return fail(UZKGE_ERR_ARG, "ntt_fr_batch: null pointer");
    if (inverse != 0 && inverse != 1) return fail(UZKGE_ERR_ARG, "ntt_fr_batch: inverse must be 0 or 1");
    API_ENTER(-1);
    if (k == 0) return UZKGE_OK;
    bool ok = false;
    ntt_root_of_unity(domain_size, &ok);
    if (!ok || domain_size % 9 == 0) return fail(UZKGE_ERR_SIZE, "ntt_fr_batch: domain size must be 2^k or 3 * 2^k, k <= 28");
    for (size_t j = 0; j < k; j++) {
        if (!inouts[j]) return fail(UZKGE_ERR_ARG, "ntt_fr_batch: null vector");
        if (len_in[j] > domain_size) return fail(UZKGE_ERR_SIZE, "ntt_fr_batch: input longer than the domain");
    }
    // Three stages on three streams over three device buffers: while transform j runs, vector j + 1 is on its way in and
    // vector j - 1 on its way out (PCIe is full duplex), so a batch costs max(H2D, D2H, compute) per vector instead of their sum.
    const size_t bytes = domain_size * sizeof(fe);
    CUDA_OR_FAIL(g.data.reserve(3 * bytes), "ntt_fr_batch: data buffers");
    CUDA_OR_FAIL(g.scratch.reserve(bytes), "ntt_fr_batch: scratch buffer");
    if (!g.copy_stream) CUDA_OR_FAIL(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking), "ntt_fr_batch: stream");
    if (!g.out_stream) CUDA_OR_FAIL(cudaStreamCreateWithFlags(&g.out_stream, cudaStreamNonBlocking), "ntt_fr_batch: stream");
    while (g.copy_events.size() < 9) {
        cudaEvent_t ev;
        CUDA_OR_FAIL(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "ntt_fr_batch: event");
        g.copy_events.push_back(ev);
    }
    cudaEvent_t* in_done = g.copy_events.data();        // [3]: vector landed in buffer b
    cudaEvent_t* run_done = g.copy_events.data() + 3;   // [3]: transform finished in buffer b
    cudaEvent_t* out_done = g.copy_events.data() + 6;   // [3]: buffer b copied back, free again
    fe shift;
    if (coset_shift) memcpy(&shift, coset_shift, sizeof(fe));
    for (size_t j = 0; j < k; j++) {
        const size_t b = j % 3;
        fe* buf = (fe*)((char*)g.data.p + b * bytes);
        if (j >= 3) CUDA_OR_FAIL(cudaStreamWaitEvent(g.copy_stream, out_done[b], 0), "ntt_fr_batch: wait");
        if (len_in[j])
            CUDA_OR_FAIL(cudaMemcpyAsync(buf, inouts[j], len_in[j] * sizeof(fe), cudaMemcpyHostToDevice, g.copy_stream), "ntt_fr_batch: H2D");
        CUDA_OR_FAIL(cudaEventRecord(in_done[b], g.copy_stream), "ntt_fr_batch: event");
        CUDA_OR_FAIL(cudaStreamWaitEvent(g.stream, in_done[b], 0), "ntt_fr_batch: wait");
        int rc = g.ntt->run(buf, buf, (fe*)g.scratch.p, len_in[j], domain_size, inverse != 0, coset_shift ? &shift : nullptr, g.stream);
        if (rc != UZKGE_OK) {
            cudaDeviceSynchronize();
            return engine_fail(rc, "ntt_fr_batch: launch");
        }
        CUDA_OR_FAIL(cudaEventRecord(run_done[b], g.stream), "ntt_fr_batch: event");
        CUDA_OR_FAIL(cudaStreamWaitEvent(g.out_stream, run_done[b], 0), "ntt_fr_batch: wait");
        CUDA_OR_FAIL(cudaMemcpyAsync(inouts[j], buf, bytes, cudaMemcpyDeviceToHost, g.out_stream), "ntt_fr_batch: D2H");
        CUDA_OR_FAIL(cudaEventRecord(out_done[b], g.out_stream), "ntt_fr_batch: event");
    }
    CUDA_OR_FAIL(cudaStreamSynchronize(g.out_stream), "ntt_fr_batch: execution");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "ntt_fr_batch: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_ntt_fr_device(const void* d_in, void* d_out, void* d_scratch, size_t len_in, size_t domain_size,
                                 int32_t inverse, const uint64_t* coset_shift_host, void* stream) {
    if (!d_in || !d_out || !d_scratch) return fail(UZKGE_ERR_ARG, "ntt_fr_device: null pointer");
    if (d_scratch == d_in || d_scratch == d_out) return fail(UZKGE_ERR_ARG, "ntt_fr_device: scratch must not alias");
    if (len_in > domain_size) return fail(UZKGE_ERR_SIZE, "ntt_fr_device: input longer than the domain");
    API_ENTER(-1);
    bool ok = false;
    ntt_root_of_unity(domain_size, &ok);
    if (!ok || domain_size % 9 == 0) return fail(UZKGE_ERR_SIZE, "ntt_fr_device: domain size must be 2^k or 3 * 2^k");
    fe shift;
    if (coset_shift_host) memcpy(&shift, coset_shift_host, sizeof(fe));
    int rc = g.ntt->run((const fe*)d_in, (fe*)d_out, (fe*)d_scratch, len_in, domain_size, inverse != 0,
                        coset_shift_host ? &shift : nullptr, (cudaStream_t)stream);
    return engine_fail(rc, "ntt_fr_device");
}

UZKGE_API int32_t uzkge_cuda_ntt_cross_fr_device(const void* d_in, void* d_out, uint32_t log_ranks, size_t cols, size_t col_offset,
                                                 size_t n_total, int32_t inverse, void* stream) {
    if (!d_in || !d_out || d_in == d_out) return fail(UZKGE_ERR_ARG, "ntt_cross_fr_device: null or aliasing pointers");
    API_ENTER(-1);
    int rc = g.ntt->cross((const fe*)d_in, (fe*)d_out, log_ranks, cols, col_offset, n_total, inverse != 0, (cudaStream_t)stream);
    return engine_fail(rc, "ntt_cross_fr_device");
}

UZKGE_API int32_t uzkge_cuda_ntt_cross_rows_fr_device(const void* const* d_in_rows, void* const* d_out_rows, uint32_t log_ranks, size_t cols,
                                                      size_t col_offset, size_t n_total, int32_t inverse, void* stream) {
    if (!d_in_rows || !d_out_rows) return fail(UZKGE_ERR_ARG, "ntt_cross_rows_fr_device: null pointer");
    API_ENTER(-1);
    int rc = g.ntt->cross_rows((const fe* const*)d_in_rows, (fe* const*)d_out_rows, log_ranks, cols, col_offset, n_total, inverse != 0,
                               (cudaStream_t)stream);
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "ntt_cross_rows_fr_device: 2, 4 or 8 ranks, power-of-two size, columns inside a slice");
    if (rc == UZKGE_ERR_ARG) return fail(rc, "ntt_cross_rows_fr_device: null row pointer");
    return engine_fail(rc, "ntt_cross_rows_fr_device");
}

UZKGE_API int32_t uzkge_cuda_dev_alloc(size_t bytes, void** d_ptr) {
    if (!d_ptr) return fail(UZKGE_ERR_ARG, "dev_alloc: null pointer");
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaMalloc(d_ptr, bytes ? bytes : 1), "dev_alloc");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_dev_free(void* d_ptr) {
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaFree(d_ptr), "dev_free");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_dev_copy_in(void* d_dst, const void* h_src, size_t bytes) {
    if (bytes && (!d_dst || !h_src)) return fail(UZKGE_ERR_ARG, "dev_copy_in: null pointer");
    API_ENTER(-1);
    if (bytes == 0) return UZKGE_OK;
    // synchronous, on the legacy default stream: ordered with *_device calls made with stream = NULL
    CUDA_OR_FAIL(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice), "dev_copy_in");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_dev_copy_out(void* h_dst, const void* d_src, size_t bytes) {
    if (bytes && (!h_dst || !d_src)) return fail(UZKGE_ERR_ARG, "dev_copy_out: null pointer");
    API_ENTER(-1);
    if (bytes == 0) return UZKGE_OK;
    // synchronous, on the legacy default stream: everything enqueued through *_device calls with stream = NULL completes first
    CUDA_OR_FAIL(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost), "dev_copy_out");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_ipc_export(const void* d_ptr, uint8_t handle[64]) {
    if (!d_ptr || !handle) return fail(UZKGE_ERR_ARG, "ipc_export: null pointer");
    API_ENTER(-1);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CUDA_OR_FAIL(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)), "ipc_export");
    memcpy(handle, &h, 64);
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_ipc_open(const uint8_t handle[64], void** d_ptr) {
    if (!d_ptr || !handle) return fail(UZKGE_ERR_ARG, "ipc_open: null pointer");
    API_ENTER(-1);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_OR_FAIL(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess), "ipc_open");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_ipc_close(void* d_ptr) {
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaIpcCloseMemHandle(d_ptr), "ipc_close");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_poly_eval_fr(const uint64_t* coefs, size_t n, const uint64_t x[4], uint64_t out[4]) {
    if (!coefs || !x || !out) return fail(UZKGE_ERR_ARG, "poly_eval_fr: null pointer");
    if (n == 0) return fail(UZKGE_ERR_SIZE, "poly_eval_fr: empty coefficient vector");
    API_ENTER(-1);
    CUDA_OR_FAIL(g.data.reserve(n * sizeof(fe)), "poly_eval_fr: buffer");
    CUDA_OR_FAIL(g.small.reserve(4096), "poly_eval_fr: buffer");
    CUDA_OR_FAIL(cudaMemcpyAsync(g.data.p, coefs, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "poly_eval_fr: H2D");
    fe z;
    memcpy(&z, x, sizeof(fe));
    int rc = g.poly->horner((const fe*)g.data.p, n, z, nullptr, (fe*)g.small.p, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "poly_eval_fr");
    CUDA_OR_FAIL(cudaMemcpyAsync(out, g.small.p, sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "poly_eval_fr: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "poly_eval_fr: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_poly_div_linear_fr(const uint64_t* coefs, size_t n, const uint64_t z_in[4], uint64_t* quotient, uint64_t rem[4]) {
    if (!coefs || !z_in || !rem || (n > 1 && !quotient)) return fail(UZKGE_ERR_ARG, "poly_div_linear_fr: null pointer");
    if (n == 0) return fail(UZKGE_ERR_SIZE, "poly_div_linear_fr: empty coefficient vector");
    API_ENTER(-1);
    CUDA_OR_FAIL(g.data.reserve(n * sizeof(fe)), "poly_div_linear_fr: buffer");
    CUDA_OR_FAIL(g.scratch.reserve(n * sizeof(fe)), "poly_div_linear_fr: buffer");
    CUDA_OR_FAIL(g.small.reserve(4096), "poly_div_linear_fr: buffer");
    CUDA_OR_FAIL(cudaMemcpyAsync(g.data.p, coefs, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "poly_div_linear_fr: H2D");
    fe z;
    memcpy(&z, z_in, sizeof(fe));
    int rc = g.poly->horner((const fe*)g.data.p, n, z, (fe*)g.scratch.p, (fe*)g.small.p, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "poly_div_linear_fr");
    if (n > 1) CUDA_OR_FAIL(cudaMemcpyAsync(quotient, g.scratch.p, (n - 1) * sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "poly_div_linear_fr: D2H");
    CUDA_OR_FAIL(cudaMemcpyAsync(rem, g.small.p, sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "poly_div_linear_fr: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "poly_div_linear_fr: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_poly_horner_fr_device(const void* d_coefs, size_t n, const uint64_t z_host[4], void* d_quotient, void* d_value,
                                                   void* stream) {
    if (!d_coefs || !z_host || !d_value) return fail(UZKGE_ERR_ARG, "poly_horner_fr_device: null pointer");
    if (n == 0) return fail(UZKGE_ERR_SIZE, "poly_horner_fr_device: empty coefficient vector");
    API_ENTER(-1);
    fe z;
    memcpy(&z, z_host, sizeof(fe));
    int rc = g.poly->horner((const fe*)d_coefs, n, z, (fe*)d_quotient, (fe*)d_value, (cudaStream_t)stream);
    return engine_fail(rc, "poly_horner_fr_device");
}

UZKGE_API int32_t uzkge_cuda_poly_eval_batch_fr_device(const void* const* d_polys, const size_t* lens, const uint32_t* point_index, size_t k,
                                                       const uint64_t* points_host, size_t npoints, void* d_values, void* stream) {
    if (k && (!d_polys || !lens || !point_index || !points_host || !d_values)) return fail(UZKGE_ERR_ARG, "poly_eval_batch_fr_device: null pointer");
    if (k > UZKGE_EVAL_BATCH_MAX || npoints > 2) return fail(UZKGE_ERR_SIZE, "poly_eval_batch_fr_device: k <= 32, at most two points");
    API_ENTER(-1);
    fe pts[2];
    memcpy(pts, points_host, npoints * sizeof(fe));
    uint64_t n64[UZKGE_EVAL_BATCH_MAX];
    for (size_t j = 0; j < k; j++) n64[j] = lens[j];
    int rc = g.poly->eval_batch((const fe* const*)d_polys, n64, point_index, (uint32_t)k, pts, (uint32_t)npoints, (fe*)d_values, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "poly_eval_batch_fr_device: empty polynomial or bad point index");
    if (rc == UZKGE_ERR_ARG) return fail(rc, "poly_eval_batch_fr_device: null polynomial");
    return engine_fail(rc, "poly_eval_batch_fr_device");
}

UZKGE_API int32_t uzkge_cuda_grand_product_fr(const uint64_t* num, const uint64_t* den, size_t n, uint64_t* out) {
    if (!out || (n && (!num || !den))) return fail(UZKGE_ERR_ARG, "grand_product_fr: null pointer");
    API_ENTER(-1);
    if (n == 0) {
        const fe one = fe_one<FrP>();
        memcpy(out, &one, sizeof(fe));
        return UZKGE_OK;
    }
    CUDA_OR_FAIL(g.data.reserve((2 * n + 1) * sizeof(fe)), "grand_product_fr: buffer");
    CUDA_OR_FAIL(g.scratch.reserve((2 * n + 2) * sizeof(fe)), "grand_product_fr: buffer");
    fe* d_num = (fe*)g.data.p;
    fe* d_den = d_num + n;
    CUDA_OR_FAIL(cudaMemcpyAsync(d_num, num, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "grand_product_fr: H2D");
    CUDA_OR_FAIL(cudaMemcpyAsync(d_den, den, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "grand_product_fr: H2D");
    // the result (n + 1 elements) overwrites the inputs' buffer once they are consumed
    int rc = g.poly->grand_product(d_num, d_den, n, d_num, (fe*)g.scratch.p, g.stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "grand_product_fr: a denominator is zero");
    if (rc != UZKGE_OK) return engine_fail(rc, "grand_product_fr");
    CUDA_OR_FAIL(cudaMemcpyAsync(out, d_num, (n + 1) * sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "grand_product_fr: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "grand_product_fr: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_plonk_quotient_fr_device(const uzkge_quotient_args* args, void* d_out, void* stream) {
    API_ENTER(-1);
    int rc = plonk_quotient_run(args, nullptr, d_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "plonk_quotient_fr_device: null pointer");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "plonk_quotient_fr_device: m must be a multiple of factor, 1 <= factor <= 16");
    return engine_fail(rc, "plonk_quotient_fr_device");
}

UZKGE_API int32_t uzkge_cuda_plonk_quotient_shuffle_fr_device(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* shuffle,
                                                              void* d_out, void* stream) {
    if (!shuffle) return fail(UZKGE_ERR_ARG, "plonk_quotient_shuffle_fr_device: null pointer");
    API_ENTER(-1);
    int rc = plonk_quotient_run(args, shuffle, d_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "plonk_quotient_shuffle_fr_device: null pointer");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "plonk_quotient_shuffle_fr_device: m must be a multiple of factor, 1 <= factor <= 16");
    return engine_fail(rc, "plonk_quotient_shuffle_fr_device");
}

UZKGE_API int32_t uzkge_cuda_fr_lincomb_device(const void* const* d_polys, const size_t* lens, const uint64_t* coefs_host, size_t k,
                                               void* d_out, size_t out_len, void* stream) {
    API_ENTER(-1);
    int rc = fr_lincomb_run(d_polys, lens, coefs_host, k, d_out, out_len, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_lincomb_device: null pointer");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "fr_lincomb_device: 1 <= k <= UZKGE_LINCOMB_MAX");
    return engine_fail(rc, "fr_lincomb_device");
}

UZKGE_API int32_t uzkge_cuda_fr_add_sparse_device(void* d_poly, const size_t* idx, const uint64_t* vals_host, size_t k, void* stream) {
    API_ENTER(-1);
    int rc = fr_add_sparse_run(d_poly, idx, vals_host, k, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_add_sparse_device: null pointer");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "fr_add_sparse_device: k <= UZKGE_SPARSE_MAX");
    return engine_fail(rc, "fr_add_sparse_device");
}

UZKGE_API int32_t uzkge_cuda_fr_powers_device(const uint64_t base_host[4], const uint64_t* scale_host, size_t n, void* d_out, void* stream) {
    API_ENTER(-1);
    int rc = fr_powers_run(base_host, scale_host, n, d_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_powers_device: null pointer");
    return engine_fail(rc, "fr_powers_device");
}

UZKGE_API int32_t uzkge_cuda_fr_gather_device(const void* d_src, const void* d_idx_u32, size_t n, void* d_out, void* stream) {
    API_ENTER(-1);
    int rc = fr_gather_run(d_src, d_idx_u32, n, d_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_gather_device: null pointer");
    return engine_fail(rc, "fr_gather_device");
}

UZKGE_API int32_t uzkge_cuda_fr_mul_device(const void* d_a, const void* d_b, size_t n, void* d_out, void* stream) {
    API_ENTER(-1);
    int rc = fr_mul_run(d_a, d_b, n, d_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_mul_device: null pointer");
    return engine_fail(rc, "fr_mul_device");
}

UZKGE_API int32_t uzkge_cuda_fr_trimmed_len_device(const void* d_poly, size_t n, size_t* len_out, void* stream) {
    API_ENTER(-1);
    CUDA_OR_FAIL(g.small.reserve(4096), "fr_trimmed_len_device: buffer");
    int rc = fr_trimmed_len_run(d_poly, n, (unsigned long long*)g.small.p, len_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_trimmed_len_device: null pointer");
    return engine_fail(rc, "fr_trimmed_len_device");
}

UZKGE_API int32_t uzkge_cuda_grand_product_fr_device(const void* d_num, const void* d_den, size_t n, void* d_out, void* d_tmp, void* stream) {
    if (!d_num || !d_den || !d_out || !d_tmp) return fail(UZKGE_ERR_ARG, "grand_product_fr_device: null pointer");
    API_ENTER(-1);
    int rc = g.poly->grand_product((const fe*)d_num, (const fe*)d_den, n, (fe*)d_out, (fe*)d_tmp, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "grand_product_fr_device: a denominator is zero");
    return engine_fail(rc, "grand_product_fr_device");
}

UZKGE_API int32_t uzkge_cuda_plonk_z_evals_fr_device(const void* const d_w[5], const void* const d_sigma[5], const void* d_group,
                                                     const uint64_t* k_host, const uint64_t beta_host[4], const uint64_t gamma_host[4], size_t n,
                                                     void* d_z, void* d_tmp, void* stream) {
    API_ENTER(-1);
    int rc = plonk_z_evals_run(g.poly.get(), d_w, d_sigma, d_group, k_host, beta_host, gamma_host, n, d_z, d_tmp, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "plonk_z_evals_fr_device: null pointer or zero denominator");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "plonk_z_evals_fr_device: n >= 2");
    return engine_fail(rc, "plonk_z_evals_fr_device");
}

UZKGE_API int32_t uzkge_cuda_fr_root_of_unity(size_t n, uint64_t out[4]) {
    if (!out) return fail(UZKGE_ERR_ARG, "fr_root_of_unity: null pointer");
    bool ok = false;
    const fe w = ntt_root_of_unity(n, &ok);
    if (!ok) return fail(UZKGE_ERR_SIZE, "fr_root_of_unity: n must be 3^a 2^b, a <= 2, b <= 28");
    memcpy(out, &w, sizeof(fe));
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_g1_add(const uint64_t a_jac[12], const uint64_t b_jac[12], uint64_t out_jac[12]) {
    if (!a_jac || !b_jac || !out_jac) return fail(UZKGE_ERR_ARG, "g1_add: null pointer");
    API_ENTER(-1);
    CUDA_OR_FAIL(g.small.reserve(4096), "g1_add: buffer");
    jacobian* d = (jacobian*)g.small.p;
    CUDA_OR_FAIL(cudaMemcpyAsync(d, a_jac, sizeof(jacobian), cudaMemcpyHostToDevice, g.stream), "g1_add: H2D");
    CUDA_OR_FAIL(cudaMemcpyAsync(d + 1, b_jac, sizeof(jacobian), cudaMemcpyHostToDevice, g.stream), "g1_add: H2D");
    int rc = g.msm->g1_add(d, d + 1, d + 2, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "g1_add");
    CUDA_OR_FAIL(cudaMemcpyAsync(out_jac, d + 2, sizeof(jacobian), cudaMemcpyDeviceToHost, g.stream), "g1_add: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "g1_add: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_g1_to_affine(const uint64_t in_jac[12], uint64_t out_affine[8]) {
    if (!in_jac || !out_affine) return fail(UZKGE_ERR_ARG, "g1_to_affine: null pointer");
    API_ENTER(-1);
    CUDA_OR_FAIL(g.small.reserve(4096), "g1_to_affine: buffer");
    jacobian* d = (jacobian*)g.small.p;
    CUDA_OR_FAIL(cudaMemcpyAsync(d, in_jac, sizeof(jacobian), cudaMemcpyHostToDevice, g.stream), "g1_to_affine: H2D");
    int rc = g.msm->g1_to_affine(d, (affine*)(d + 1), g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "g1_to_affine");
    CUDA_OR_FAIL(cudaMemcpyAsync(out_affine, d + 1, sizeof(affine), cudaMemcpyDeviceToHost, g.stream), "g1_to_affine: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "g1_to_affine: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_srs_generate(const uint64_t tau[4], size_t n, uint64_t* out_affine_xy) {
    if (!tau || (n && !out_affine_xy)) return fail(UZKGE_ERR_ARG, "srs_generate: null pointer");
    if (n >= (1ull << 28)) return fail(UZKGE_ERR_SIZE, "srs_generate: n too large");
    API_ENTER(-1);
    fe t;
    memcpy(&t, tau, sizeof(fe));
    const size_t slab = 1u << 22;
    CUDA_OR_FAIL(g.data.reserve((n < slab ? n : slab) * sizeof(affine) + 64), "srs_generate: buffer");
    for (size_t first = 0; first < n; first += slab) {
        const uint32_t count = (uint32_t)((n - first) < slab ? (n - first) : slab);
        int rc = g.msm->powers_of_tau(t, first, count, (affine*)g.data.p, g.stream);
        if (rc != UZKGE_OK) return engine_fail(rc, "srs_generate: launch");
        CUDA_OR_FAIL(cudaMemcpyAsync(out_affine_xy + first * 8, g.data.p, count * sizeof(affine), cudaMemcpyDeviceToHost, g.stream),
                     "srs_generate: D2H");
        CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "srs_generate: execution");
    }
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_srs_generate_lagrange(const uint64_t tau[4], size_t n, uint64_t* out_affine_xy) {
    if (!tau || !out_affine_xy) return fail(UZKGE_ERR_ARG, "srs_generate_lagrange: null pointer");
    if (n == 0 || (n & (n - 1)) || n > (1ull << 26)) return fail(UZKGE_ERR_SIZE, "srs_generate_lagrange: n must be a power of two <= 2^26");
    API_ENTER(-1);
    CUDA_OR_FAIL(g.data.reserve(n * sizeof(affine) + 64), "srs_generate_lagrange: buffer");
    CUDA_OR_FAIL(g.scratch.reserve(2 * n * sizeof(fe) + 64), "srs_generate_lagrange: buffer");
    fe* pw = (fe*)g.scratch.p;
    fe* tmp = pw + n;
    int rc = fr_powers_run(tau, nullptr, n, pw, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "srs_generate_lagrange: powers");
    rc = g.ntt->run(pw, pw, tmp, n, n, true, nullptr, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "srs_generate_lagrange: transform");
    rc = g.msm->fixed_base_mul(pw, (uint32_t)n, (affine*)g.data.p, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "srs_generate_lagrange: launch");
    CUDA_OR_FAIL(cudaMemcpyAsync(out_affine_xy, g.data.p, n * sizeof(affine), cudaMemcpyDeviceToHost, g.stream), "srs_generate_lagrange: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "srs_generate_lagrange: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_srs_lagrange_from_monomial(const uint64_t* monomial_affine_xy, size_t n, uint64_t* out_affine_xy) {
    if (!monomial_affine_xy || !out_affine_xy) return fail(UZKGE_ERR_ARG, "srs_lagrange_from_monomial: null pointer");
    if (n == 0 || (n & (n - 1)) || n > (1ull << 24)) return fail(UZKGE_ERR_SIZE, "srs_lagrange_from_monomial: n must be a power of two <= 2^24");
    API_ENTER(-1);
    uint32_t log_n = 0;
    while ((1ull << log_n) < n) log_n++;
    bool ok = false;
    const fe w = ntt_root_of_unity(n, &ok);
    if (!ok) return fail(UZKGE_ERR_SIZE, "srs_lagrange_from_monomial: no domain of this size");
    const fe w_inv = fe_inv<FrP>(w);
    fe nf = fe_zero();
    nf.l[0] = (uint32_t)n;
    const fe n_inv = fe_from_mont<FrP>(fe_inv<FrP>(fe_to_mont<FrP>(nf)));   // canonical bits for the double-and-add
    CUDA_OR_FAIL(g.data.reserve(n * sizeof(affine) + 64), "srs_lagrange_from_monomial: buffer");
    CUDA_OR_FAIL(g.scratch.reserve(n * sizeof(xyzz) + (n / 2 + 1) * sizeof(fe) + 64), "srs_lagrange_from_monomial: buffer");
    affine* d_pts = (affine*)g.data.p;
    xyzz* d_work = (xyzz*)g.scratch.p;
    fe* d_tw = (fe*)(d_work + n);
    CUDA_OR_FAIL(cudaMemcpyAsync(d_pts, monomial_affine_xy, n * sizeof(affine), cudaMemcpyHostToDevice, g.stream), "srs_lagrange_from_monomial: H2D");
    int rc = fr_powers_run((const uint64_t*)&w_inv, nullptr, n / 2 ? n / 2 : 1, d_tw, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "srs_lagrange_from_monomial: twiddles");
    rc = ec_intt_run(d_pts, (uint32_t)n, log_n, d_tw, n_inv, d_work, d_pts, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "srs_lagrange_from_monomial: launch");
    CUDA_OR_FAIL(cudaMemcpyAsync(out_affine_xy, d_pts, n * sizeof(affine), cudaMemcpyDeviceToHost, g.stream), "srs_lagrange_from_monomial: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "srs_lagrange_from_monomial: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_msm_g1_small_device(uint64_t handle, const size_t* idx, const uint64_t* scalars_host, size_t k, int32_t accumulate,
                                                 void* d_out_jac, void* stream) {
    if (!d_out_jac || (k && (!idx || !scalars_host))) return fail(UZKGE_ERR_ARG, "msm_g1_small_device: null pointer");
    API_ENTER(-1);
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1_small_device: unknown handle");
    int rc = g.msm->small_msm(&it->second, idx, scalars_host, (uint32_t)k, accumulate != 0, (jacobian*)d_out_jac, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "msm_g1_small_device: k <= 32, indices inside the SRS");
    return engine_fail(rc, "msm_g1_small_device");
}

UZKGE_API int32_t uzkge_cuda_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(UZKGE_ERR_ARG, "host_alloc: null pointer");
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault), "host_alloc");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_host_free(void* p) {
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaFreeHost(p), "host_free");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_host_register(void* p, size_t bytes) {
    if (!p) return fail(UZKGE_ERR_ARG, "host_register: null pointer");
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaHostRegister(p, bytes, cudaHostRegisterDefault), "host_register");
    return UZKGE_OK;
}
UZKGE_API int32_t uzkge_cuda_host_unregister(void* p) {
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaHostUnregister(p), "host_unregister");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_profile_enable(int32_t on) {
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaDeviceSynchronize(), "profile_enable");
    double sums[Profiler::KINDS][Profiler::MAX_PHASES] = {};
    uint64_t runs[Profiler::KINDS] = {};
    g_prof.collect(sums, runs);  // drop stale records
    g_prof.enabled = on != 0;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_profile_read(int32_t kind, double phase_ms[8], uint64_t* runs_out) {
    if (!phase_ms || !runs_out) return fail(UZKGE_ERR_ARG, "profile_read: null pointer");
    if (kind < 0 || kind >= Profiler::KINDS) return fail(UZKGE_ERR_ARG, "profile_read: kind must be 0 (MSM) or 1 (NTT)");
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaDeviceSynchronize(), "profile_read");
    static double sums[Profiler::KINDS][Profiler::MAX_PHASES];
    static uint64_t runs[Profiler::KINDS];
    g_prof.collect(sums, runs);
    for (int i = 0; i < Profiler::MAX_PHASES; i++) {
        phase_ms[i] = sums[kind][i];
        sums[kind][i] = 0;
    }
    *runs_out = runs[kind];
    runs[kind] = 0;
    return UZKGE_OK;
}

UZKGE_API uint64_t uzkge_cuda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

UZKGE_API int32_t uzkge_cuda_configure(const char* key, uint64_t value) {
    if (!key) return fail(UZKGE_ERR_ARG, "configure: null key");
    std::lock_guard<std::mutex> lock(g.mu);
    const std::string k(key);
    if (k == "msm_lanes") {
        if (value > 32 || (value & (value - 1))) return fail(UZKGE_ERR_ARG, "configure: msm_lanes must be 0 or a power of two <= 32");
        int rc = ensure_init(-1);
        if (rc != UZKGE_OK) return rc;
        g.msm->force_lanes((uint32_t)value);
        return UZKGE_OK;
    }
    if (k == "ntt_big_threads") {
        if (value != 512 && value != 1024) return fail(UZKGE_ERR_ARG, "configure: ntt_big_threads is 512 or 1024");
        int rc = ensure_init(-1);
        if (rc != UZKGE_OK) return rc;
        g.ntt->set_big_threads((uint32_t)value);
        return UZKGE_OK;
    }
    if (k == "ntt_log_tile" || k == "ntt_max_log_r" || k == "ntt_two_pass_max") {
        if (k == "ntt_log_tile") {
            if (value < 4 || value > 12) return fail(UZKGE_ERR_ARG, "configure: ntt_log_tile in 4..12");
            g.ntt_log_tile = (uint32_t)value;
        } else if (k == "ntt_max_log_r") {
            if (value < 4 || value > 12) return fail(UZKGE_ERR_ARG, "configure: ntt_max_log_r in 4..12");
            g.ntt_max_log_r = (uint32_t)value;
        } else {
            g.ntt_two_pass_max = (uint32_t)value;
        }
        if (g.ready) {  // plans are cached per size: start over with the new limits
            cudaStreamSynchronize(g.stream);
            g.ntt.reset(new NttEngine(g.sm_count));
            g.ntt->configure(g.ntt_log_tile, g.ntt_max_log_r, g.ntt_two_pass_max);
        }
        return UZKGE_OK;
    }
    if (k == "msm_affine") {
        int rc = ensure_init(-1);
        if (rc != UZKGE_OK) return rc;
        g.msm->set_affine((uint32_t)value);
        return UZKGE_OK;
    }
    if (k == "quotient_min_blocks") {
        g_quotient_min_blocks = (int)value;
        return UZKGE_OK;
    }
    return fail(UZKGE_ERR_ARG, "configure: unknown key");
}

UZKGE_API int32_t uzkge_cuda_field_mul(int32_t field, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    if (n && (!a || !b || !out)) return fail(UZKGE_ERR_ARG, "field_mul: null pointer");
    if (field != 0 && field != 1) return fail(UZKGE_ERR_ARG, "field_mul: field must be 0 (Fr) or 1 (Fq)");
    API_ENTER(-1);
    if (n == 0) return UZKGE_OK;
    CUDA_OR_FAIL(g.data.reserve(n * sizeof(fe)), "field_mul: buffer");
    CUDA_OR_FAIL(g.scratch.reserve(n * sizeof(fe)), "field_mul: buffer");
    CUDA_OR_FAIL(cudaMemcpyAsync(g.data.p, a, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "field_mul: H2D");
    CUDA_OR_FAIL(cudaMemcpyAsync(g.scratch.p, b, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "field_mul: H2D");
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (field == 0)
        field_mul_kernel<FrP><<<grid, 256, 0, g.stream>>>((const fe*)g.data.p, (const fe*)g.scratch.p, (fe*)g.data.p, n);
    else
        field_mul_kernel<FqP><<<grid, 256, 0, g.stream>>>((const fe*)g.data.p, (const fe*)g.scratch.p, (fe*)g.data.p, n);
    UZ_COUNT_LAUNCH(1);
    CUDA_OR_FAIL(cudaGetLastError(), "field_mul: launch");
    CUDA_OR_FAIL(cudaMemcpyAsync(out, g.data.p, n * sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "field_mul: D2H");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "field_mul: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_bench_field_mul(int32_t field, uint32_t iters, double* muls_per_s) {
    if (!muls_per_s) return fail(UZKGE_ERR_ARG, "bench_field_mul: null pointer");
    if (field != 0 && field != 1) return fail(UZKGE_ERR_ARG, "bench_field_mul: field must be 0 (Fr) or 1 (Fq)");
    API_ENTER(-1);
    const unsigned grid = (unsigned)g.sm_count * 8, nt = 256;
    CUDA_OR_FAIL(g.data.reserve((size_t)grid * nt * sizeof(fe)), "bench_field_mul: buffer");
    cudaEvent_t e0, e1;
    CUDA_OR_FAIL(cudaEventCreate(&e0), "event");
    CUDA_OR_FAIL(cudaEventCreate(&e1), "event");
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, g.stream);
        if (field == 0)
            field_mul_bench_kernel<FrP><<<grid, nt, 0, g.stream>>>((fe*)g.data.p, iters);
        else
            field_mul_bench_kernel<FqP><<<grid, nt, 0, g.stream>>>((fe*)g.data.p, iters);
        UZ_COUNT_LAUNCH(1);
        cudaEventRecord(e1, g.stream);
        cudaError_t e = cudaStreamSynchronize(g.stream);
        if (e != cudaSuccess) return fail_cuda("bench_field_mul: execution", e);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *muls_per_s = (double)grid * nt * 4.0 * iters / (best * 1e-3);
    return UZKGE_OK;
}

}  // extern "C"
