// api.cu -- the C ABI of include/uzkge_cuda.h: process-wide state, host<->device staging, error mapping.
//
// The Rust side binds exactly these symbols from the bodies of
// KZGCommitmentSchemeBN254::commit (/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:278-293) and
// FpPolynomial::{fft,ifft,coset_fft,coset_ifft}_with_domain (/root/reference/uzkge/src/poly_commit/field_polynomial.rs:583-607).
// There is no CPU path in this library: without a Blackwell device every call fails.
#include <cuda_runtime.h>

#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

#include "devmem.cuh"
#include "internal.h"

namespace uz {
std::atomic<uint64_t> g_launches{0};
static Profiler g_profs[UZ_MAX_DEVICES];
Profiler& prof_for_current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return g_profs[(d >= 0 && d < UZ_MAX_DEVICES) ? d : 0];
}
}

using namespace uz;

namespace uz {
extern int g_quotient_min_blocks;  // quotient.cu
extern int g_group_deal_min_log_n; // prover.cu
static int g_virtual_devices = 0;   // tests: group members cycle over the visible GPUs (uzkge_cuda_configure "virtual_devices")
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

// Orders the users of a workspace that is shared by every stream of a device (the MSM buffers of an SRS handle, the scan
// workspace of the polynomial engine, the small result buffer): the next user's stream waits for the previous user's last
// launch when the two streams differ.  Same stream: nothing is enqueued.
struct Fence {
    cudaEvent_t ev = nullptr;
    cudaStream_t last = nullptr;
    bool used = false;
    cudaError_t enter(cudaStream_t st) {
        if (used && last != st) return cudaStreamWaitEvent(st, ev, 0);
        return cudaSuccess;
    }
    cudaError_t leave(cudaStream_t st) {
        if (!ev) {
            cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        last = st;
        used = true;
        return cudaEventRecord(ev, st);
    }
};

// One per CUDA device, created on first use.  Every entry point works on ONE device's state under that device's mutex: calls
// aimed at different devices (different threads, or the workers of a multi-device call) run concurrently.
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
    std::map<uint64_t, Fence> srs_fence;
    Fence poly_fence, small_fence;
    uint64_t next_handle = 1;
    DevBuf data, scratch, scalars, small;
    cudaStream_t copy_stream = nullptr, out_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
};
struct Config {
    uint32_t ntt_log_tile = 10, ntt_max_log_r = 10, ntt_two_pass_max = 18;  // measured best on B200 (scripts/gpu_ntt_cfg.py)
    uint32_t ntt_big_threads = 1024, msm_lanes = 0, msm_affine = 0, ntt_radix4 = 19;
};
std::mutex g_reg_mu;                  // guards the registry below (never held while a device's work is enqueued)
State* g_states[UZ_MAX_DEVICES] = {};
Config g_cfg;
int g_default_device = -1;            // the first device initialised: where calls of threads that never chose a device go
thread_local int t_device = -1;       // the device this thread's calls go to (uzkge_cuda_init / uzkge_cuda_set_device)

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

}  // namespace
namespace uz {
int api_fail(int code, const char* what) { return fail(code, what); }   // for the other translation units (prover.cu)
}
namespace {

// SRS handles carry their device: bits 48..55 = device + 1, bit 63 = multi-device handle (see the bottom of this file)
constexpr uint64_t HANDLE_MULTI = 1ull << 63;
inline int handle_device(uint64_t h) { return (int)((h >> 48) & 0xff) - 1; }

// Which device does this call go to?  `want` >= 0: that device; -1: the calling thread's device, else the process default, else
// the device current in the CUDA runtime (e.g. the one torch selected).  Creates and initialises the state on first use.
int enter_state(int want, State** out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(UZKGE_ERR_NO_DEVICE, "no CUDA device: uzkge-b200 has no CPU path");
    }
    int dev = want >= 0 ? want : t_device;
    if (dev < 0) dev = g_default_device;
    if (dev < 0) {
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return fail_cuda("cudaGetDevice", e);
    }
    if (dev >= count || dev >= UZ_MAX_DEVICES) return fail(UZKGE_ERR_NO_DEVICE, "device index out of range");
    std::lock_guard<std::mutex> reg(g_reg_mu);
    State* st = g_states[dev];
    if (!st) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return fail_cuda("cudaGetDeviceProperties", e);
        if (prop.major != 10) {
            char buf[256];
            snprintf(buf, sizeof buf, "device %d (%.64s, sm_%d%d) is not a Blackwell sm_100 part: kernels are built for sm_100a only",
                     dev, prop.name, prop.major, prop.minor);
            return fail(UZKGE_ERR_NO_DEVICE, buf);
        }
        e = cudaSetDevice(dev);
        if (e != cudaSuccess) return fail_cuda("cudaSetDevice", e);
        std::unique_ptr<State> fresh(new State());
        e = cudaStreamCreateWithFlags(&fresh->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) return fail_cuda("cudaStreamCreate", e);
        fresh->device = dev;
        fresh->sm_count = prop.multiProcessorCount;
        fresh->ntt.reset(new NttEngine(fresh->sm_count));
        fresh->ntt->configure(g_cfg.ntt_log_tile, g_cfg.ntt_max_log_r, g_cfg.ntt_two_pass_max);
        fresh->ntt->set_big_threads(g_cfg.ntt_big_threads);
        fresh->ntt->set_radix4(g_cfg.ntt_radix4);
        fresh->msm.reset(new MsmEngine(fresh->sm_count));
        fresh->msm->force_lanes(g_cfg.msm_lanes);
        fresh->msm->set_affine(g_cfg.msm_affine);
        fresh->poly.reset(new PolyEngine());
        fresh->ready = true;
        st = g_states[dev] = fresh.release();
        if (g_default_device < 0) g_default_device = dev;
    }
    *out = st;
    return UZKGE_OK;
}

// st__ is the device state of this call; `g` reads like the single global it used to be
#define g (*st__)
#define API_ENTER(dev)                                                   \
    State* st__ = nullptr;                                               \
    do {                                                                 \
        int rc__ = enter_state((dev), &st__);                            \
        if (rc__ != UZKGE_OK) return rc__;                               \
    } while (0);                                                         \
    std::lock_guard<std::mutex> lock__(st__->mu);                        \
    do {                                                                 \
        cudaError_t e__ = cudaSetDevice(st__->device);                   \
        if (e__ != cudaSuccess) return fail_cuda("cudaSetDevice", e__);  \
    } while (0)
// entry points that take an SRS handle run on the handle's device, whatever device the thread selected
#define API_ENTER_HANDLE(handle, what)                                                        \
    if (handle_device(handle) < 0) return fail(UZKGE_ERR_HANDLE, what ": unknown handle");    \
    API_ENTER(handle_device(handle))

#define CUDA_OR_FAIL(expr, what)                         \
    do {                                                 \
        cudaError_t e__ = (expr);                        \
        if (e__ != cudaSuccess) return fail_cuda(what, e__); \
    } while (0)

// run `expr` (an engine call that enqueues work on `st` using a workspace guarded by `f`) ordered after the previous user of that
// workspace on any other stream
#define FENCED(f, st, what, expr)                                                   \
    do {                                                                            \
        CUDA_OR_FAIL((f).enter(st), what ": stream order");                         \
        rc = (expr);                                                                \
        cudaError_t le__ = (f).leave(st);                                           \
        if (rc == UZKGE_OK && le__ != cudaSuccess) return fail_cuda(what ": stream order", le__); \
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
    t_device = st__->device;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_set_device(int32_t device) {
    if (device < 0) return fail(UZKGE_ERR_ARG, "set_device: device index must be >= 0");
    API_ENTER(device);
    t_device = st__->device;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_get_device(void) {
    if (t_device >= 0) return t_device;
    return g_default_device;
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

static int32_t srs_upload_one(int device, const uint64_t* affine_xy, size_t n, uint32_t window_bits, uint64_t* handle) {
    if (!affine_xy || !handle) return fail(UZKGE_ERR_ARG, "srs_upload: null pointer");
    API_ENTER(device);
    MsmSrs s;
    int rc = g.msm->upload(affine_xy, n, window_bits, &s, g.stream);
    if (rc != UZKGE_OK) {
        g.msm->release(&s);
        return engine_fail(rc, "srs_upload");
    }
    const uint64_t h = ((uint64_t)(g.device + 1) << 48) | g.next_handle++;
    g.srs[h] = s;
    *handle = h;
    return UZKGE_OK;
}

static int32_t srs_free_one(uint64_t handle) {
    API_ENTER_HANDLE(handle, "srs_free");
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "srs_free: unknown handle");
    cudaStreamSynchronize(g.stream);
    cudaDeviceSynchronize();   // *_device calls may still be running on the caller's streams
    g.msm->release(&it->second);
    g.srs.erase(it);
    auto fit = g.srs_fence.find(handle);
    if (fit != g.srs_fence.end()) {
        if (fit->second.ev) cudaEventDestroy(fit->second.ev);
        g.srs_fence.erase(fit);
    }
    return UZKGE_OK;
}

static int32_t srs_info_one(uint64_t handle, uzkge_srs_info* info) {
    if (!info) return fail(UZKGE_ERR_ARG, "srs_info: null pointer");
    API_ENTER_HANDLE(handle, "srs_info");
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

static int32_t msm_g1_batch_one(uint64_t handle, const uint64_t* const* scalars, const size_t* n, size_t k, uint64_t* out_jac) {
    if (k && (!scalars || !n || !out_jac)) return fail(UZKGE_ERR_ARG, "msm_g1_batch: null pointer");
    API_ENTER_HANDLE(handle, "msm_g1");
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
    int rc;
    CUDA_OR_FAIL(g.small_fence.enter(g.stream), "msm_g1_batch: stream order");
    FENCED(g.srs_fence[handle], g.stream, "msm_g1_batch", g.msm->run_pipelined(&s, 0, ptrs.data(), n, k, d_o, g.stream, g.copy_events.data()));
    if (rc != UZKGE_OK) {
        cudaStreamSynchronize(g.copy_stream);
        cudaStreamSynchronize(g.stream);
        return engine_fail(rc, "msm_g1_batch: launch");
    }
    CUDA_OR_FAIL(cudaMemcpyAsync(out_jac, d_o, k * sizeof(jacobian), cudaMemcpyDeviceToHost, g.stream), "msm_g1: D2H");
    CUDA_OR_FAIL(g.small_fence.leave(g.stream), "msm_g1_batch: stream order");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "msm_g1: execution");
    return UZKGE_OK;
}

static int32_t msm_g1_one(uint64_t handle, size_t base_offset, const uint64_t* scalars, size_t n, uint64_t out_jac[12]) {
    if (!out_jac || (n && !scalars)) return fail(UZKGE_ERR_ARG, "msm_g1: null pointer");
    API_ENTER_HANDLE(handle, "msm_g1");
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1: unknown handle");
    MsmSrs& s = it->second;
    if (base_offset > s.n || n > s.n - base_offset) return fail(UZKGE_ERR_SIZE, "msm_g1: range outside the SRS");
    CUDA_OR_FAIL(g.scalars.reserve(n * sizeof(fe) + 32), "msm_g1: scalar buffer");
    CUDA_OR_FAIL(g.small.reserve(4096), "msm_g1: output buffer");
    if (n) CUDA_OR_FAIL(cudaMemcpyAsync(g.scalars.p, scalars, n * sizeof(fe), cudaMemcpyHostToDevice, g.stream), "msm_g1: H2D");
    int rc;
    CUDA_OR_FAIL(g.small_fence.enter(g.stream), "msm_g1: stream order");
    FENCED(g.srs_fence[handle], g.stream, "msm_g1", g.msm->run(&s, base_offset, (const fe*)g.scalars.p, n, (jacobian*)g.small.p, g.stream));
    if (rc != UZKGE_OK) {
        cudaStreamSynchronize(g.stream);
        return engine_fail(rc, "msm_g1: launch");
    }
    CUDA_OR_FAIL(cudaMemcpyAsync(out_jac, g.small.p, sizeof(jacobian), cudaMemcpyDeviceToHost, g.stream), "msm_g1: D2H");
    CUDA_OR_FAIL(g.small_fence.leave(g.stream), "msm_g1: stream order");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "msm_g1: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_msm_g1_device(uint64_t handle, size_t base_offset, const void* d_scalars, size_t n, void* d_out_jac, void* stream) {
    if (!d_out_jac || (n && !d_scalars)) return fail(UZKGE_ERR_ARG, "msm_g1_device: null pointer");
    API_ENTER_HANDLE(handle, "msm_g1_device");
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1_device: unknown handle");
    int rc;
    FENCED(g.srs_fence[handle], (cudaStream_t)stream, "msm_g1_device",
           g.msm->run(&it->second, base_offset, (const fe*)d_scalars, n, (jacobian*)d_out_jac, (cudaStream_t)stream));
    return engine_fail(rc, "msm_g1_device");
}

UZKGE_API int32_t uzkge_cuda_msm_g1_batch_device(uint64_t handle, size_t base_offset, const void* const* d_scalars, const size_t* n,
                                                 size_t k, void* d_out_jac, void* stream) {
    if (k && (!d_scalars || !n || !d_out_jac)) return fail(UZKGE_ERR_ARG, "msm_g1_batch_device: null pointer");
    API_ENTER_HANDLE(handle, "msm_g1_batch_device");
    auto it = g.srs.find(handle);
    if (it == g.srs.end()) return fail(UZKGE_ERR_HANDLE, "msm_g1_batch_device: unknown handle");
    int rc;
    FENCED(g.srs_fence[handle], (cudaStream_t)stream, "msm_g1_batch_device",
           g.msm->run_pipelined(&it->second, base_offset, (const fe* const*)d_scalars, n, k, (jacobian*)d_out_jac, (cudaStream_t)stream));
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

UZKGE_API int32_t uzkge_cuda_ntt_fr_batch_device(const void* const* d_ins, void* const* d_outs, void* d_scratch, const size_t* len_in, size_t k,
                                                 size_t domain_size, int32_t inverse, const uint64_t* coset_shift_host, void* stream) {
    if (k == 0) return UZKGE_OK;
    if (!d_ins || !d_outs || !d_scratch || !len_in) return fail(UZKGE_ERR_ARG, "ntt_fr_batch_device: null pointer");
    if (k > NTT_MAX_BATCH) return fail(UZKGE_ERR_SIZE, "ntt_fr_batch_device: at most 16 vectors per call");
    API_ENTER(-1);
    bool ok = false;
    ntt_root_of_unity(domain_size, &ok);
    if (!ok || domain_size % 9 == 0) return fail(UZKGE_ERR_SIZE, "ntt_fr_batch_device: domain size must be 2^k or 3 * 2^k");
    uint64_t lens[NTT_MAX_BATCH];
    for (size_t j = 0; j < k; j++) {
        if (!d_ins[j] || !d_outs[j]) return fail(UZKGE_ERR_ARG, "ntt_fr_batch_device: null vector");
        if (len_in[j] > domain_size) return fail(UZKGE_ERR_SIZE, "ntt_fr_batch_device: input longer than the domain");
        lens[j] = len_in[j];
    }
    fe shift;
    if (coset_shift_host) memcpy(&shift, coset_shift_host, sizeof(fe));
    int rc = g.ntt->run_batch((const fe* const*)d_ins, (fe* const*)d_outs, (fe*)d_scratch, lens, (uint32_t)k, domain_size, inverse != 0,
                              coset_shift_host ? &shift : nullptr, (cudaStream_t)stream);
    return engine_fail(rc, "ntt_fr_batch_device");
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

UZKGE_API int32_t uzkge_cuda_ntt_fr_scatter_device(const void* d_in, void* const* d_out_rows, void* d_scratch, size_t n, int32_t inverse,
                                                   uint32_t log_ranks, uint32_t rank, void* stream) {
    if (!d_in || !d_out_rows || !d_scratch) return fail(UZKGE_ERR_ARG, "ntt_fr_scatter_device: null pointer");
    if (log_ranks < 1 || log_ranks > 3 || rank >= (1u << log_ranks)) return fail(UZKGE_ERR_SIZE, "ntt_fr_scatter_device: 2, 4 or 8 ranks");
    API_ENTER(-1);
    NttScatter sc;
    for (uint32_t r = 0; r < 8; r++) {
        sc.rows[r] = r < (1u << log_ranks) ? (fe*)d_out_rows[r] : nullptr;
        if (r < (1u << log_ranks) && !sc.rows[r]) return fail(UZKGE_ERR_ARG, "ntt_fr_scatter_device: null row pointer");
    }
    sc.log_g = log_ranks;
    sc.k1 = rank;
    const fe* in = (const fe*)d_in;
    fe* out = (fe*)d_scratch;             // unused by the scattering last pass (the earlier passes work in the scratch vector)
    const uint64_t len = n;
    int rc = g.ntt->run_batch(&in, &out, (fe*)d_scratch, &len, 1, n, inverse != 0, nullptr, (cudaStream_t)stream, &sc);
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "ntt_fr_scatter_device: n must be a power of two, at least the number of ranks");
    return engine_fail(rc, "ntt_fr_scatter_device");
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
    int rc;
    CUDA_OR_FAIL(g.small_fence.enter(g.stream), "poly_eval_fr: stream order");
    FENCED(g.poly_fence, g.stream, "poly_eval_fr", g.poly->horner((const fe*)g.data.p, n, z, nullptr, (fe*)g.small.p, g.stream));
    if (rc != UZKGE_OK) return engine_fail(rc, "poly_eval_fr");
    CUDA_OR_FAIL(cudaMemcpyAsync(out, g.small.p, sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "poly_eval_fr: D2H");
    CUDA_OR_FAIL(g.small_fence.leave(g.stream), "poly_eval_fr: stream order");
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
    int rc;
    CUDA_OR_FAIL(g.small_fence.enter(g.stream), "poly_div_linear_fr: stream order");
    FENCED(g.poly_fence, g.stream, "poly_div_linear_fr", g.poly->horner((const fe*)g.data.p, n, z, (fe*)g.scratch.p, (fe*)g.small.p, g.stream));
    if (rc != UZKGE_OK) return engine_fail(rc, "poly_div_linear_fr");
    if (n > 1) CUDA_OR_FAIL(cudaMemcpyAsync(quotient, g.scratch.p, (n - 1) * sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "poly_div_linear_fr: D2H");
    CUDA_OR_FAIL(cudaMemcpyAsync(rem, g.small.p, sizeof(fe), cudaMemcpyDeviceToHost, g.stream), "poly_div_linear_fr: D2H");
    CUDA_OR_FAIL(g.small_fence.leave(g.stream), "poly_div_linear_fr: stream order");
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
    int rc;
    FENCED(g.poly_fence, (cudaStream_t)stream, "poly_horner_fr_device", g.poly->horner((const fe*)d_coefs, n, z, (fe*)d_quotient, (fe*)d_value, (cudaStream_t)stream));
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
    int rc;
    FENCED(g.poly_fence, (cudaStream_t)stream, "poly_eval_batch_fr_device",
           g.poly->eval_batch((const fe* const*)d_polys, n64, point_index, (uint32_t)k, pts, (uint32_t)npoints, (fe*)d_values, (cudaStream_t)stream));
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
    int rc;
    FENCED(g.poly_fence, g.stream, "grand_product_fr", g.poly->grand_product(d_num, d_den, n, d_num, (fe*)g.scratch.p, g.stream));
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

UZKGE_API int32_t uzkge_cuda_plonk_quotient_range_fr_device(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* shuffle,
                                                            uint64_t start, uint64_t step, uint64_t count, void* d_out, void* stream) {
    API_ENTER(-1);
    int rc = plonk_quotient_range_run(args, shuffle, start, step, count, d_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "plonk_quotient_range_fr_device: null pointer");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "plonk_quotient_range_fr_device: m a multiple of factor <= 16, start + step * (count - 1) < m");
    return engine_fail(rc, "plonk_quotient_range_fr_device");
}

UZKGE_API int32_t uzkge_cuda_plonk_coset_combine_fr_device(const void* d_u, size_t n, size_t factor, const uint64_t k1_host[4], void* d_out, void* stream) {
    API_ENTER(-1);
    int rc = plonk_coset_combine_run(d_u, n, factor, k1_host, d_out, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "plonk_coset_combine_fr_device: null pointer");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "plonk_coset_combine_fr_device: n a power of two, factor <= 16, factor * n a supported domain");
    return engine_fail(rc, "plonk_coset_combine_fr_device");
}

UZKGE_API int32_t uzkge_cuda_fr_strided_copy_device(const void* d_src, size_t src_start, size_t src_step, void* d_dst, size_t dst_start,
                                                    size_t dst_step, size_t count, void* stream) {
    API_ENTER(-1);
    int rc = fr_strided_copy_run(d_src, src_start, src_step, d_dst, dst_start, dst_step, count, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_strided_copy_device: null pointer");
    return engine_fail(rc, "fr_strided_copy_device");
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

UZKGE_API int32_t uzkge_cuda_fr_add_sparse_multi_device(void* const* d_polys, const size_t* idx, const uint64_t* vals_host, size_t k, void* stream) {
    API_ENTER(-1);
    int rc = fr_add_sparse_multi_run(d_polys, idx, vals_host, k, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_add_sparse_multi_device: null pointer");
    if (rc == UZKGE_ERR_SIZE) return fail(rc, "fr_add_sparse_multi_device: k <= UZKGE_SPARSE_MULTI_MAX");
    return engine_fail(rc, "fr_add_sparse_multi_device");
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

UZKGE_API int32_t uzkge_cuda_fr_gather_scatter_device(const void* d_src, const void* d_src_idx_u32, void* d_dst, const void* d_dst_idx_u32, size_t k,
                                                      void* stream) {
    API_ENTER(-1);
    int rc = fr_gather_scatter_run(d_src, d_src_idx_u32, d_dst, d_dst_idx_u32, k, (cudaStream_t)stream);
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_gather_scatter_device: null pointer");
    return engine_fail(rc, "fr_gather_scatter_device");
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
    int rc;
    FENCED(g.small_fence, (cudaStream_t)stream, "fr_trimmed_len_device",
           fr_trimmed_len_run(d_poly, n, (unsigned long long*)g.small.p, len_out, (cudaStream_t)stream));
    if (rc == UZKGE_ERR_ARG) return fail(rc, "fr_trimmed_len_device: null pointer");
    return engine_fail(rc, "fr_trimmed_len_device");
}

UZKGE_API int32_t uzkge_cuda_grand_product_fr_device(const void* d_num, const void* d_den, size_t n, void* d_out, void* d_tmp, void* stream) {
    if (!d_num || !d_den || !d_out || !d_tmp) return fail(UZKGE_ERR_ARG, "grand_product_fr_device: null pointer");
    API_ENTER(-1);
    int rc;
    FENCED(g.poly_fence, (cudaStream_t)stream, "grand_product_fr_device",
           g.poly->grand_product((const fe*)d_num, (const fe*)d_den, n, (fe*)d_out, (fe*)d_tmp, (cudaStream_t)stream));
    if (rc == UZKGE_ERR_ARG) return fail(rc, "grand_product_fr_device: a denominator is zero");
    return engine_fail(rc, "grand_product_fr_device");
}

UZKGE_API int32_t uzkge_cuda_plonk_z_evals_fr_device(const void* const d_w[5], const void* const d_sigma[5], const void* d_group,
                                                     const uint64_t* k_host, const uint64_t beta_host[4], const uint64_t gamma_host[4], size_t n,
                                                     void* d_z, void* d_tmp, void* stream) {
    API_ENTER(-1);
    int rc;
    FENCED(g.poly_fence, (cudaStream_t)stream, "plonk_z_evals_fr_device",
           plonk_z_evals_run(g.poly.get(), d_w, d_sigma, d_group, k_host, beta_host, gamma_host, n, d_z, d_tmp, (cudaStream_t)stream));
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
    CUDA_OR_FAIL(g.small_fence.enter(g.stream), "g1_add: stream order");
    CUDA_OR_FAIL(cudaMemcpyAsync(d, a_jac, sizeof(jacobian), cudaMemcpyHostToDevice, g.stream), "g1_add: H2D");
    CUDA_OR_FAIL(cudaMemcpyAsync(d + 1, b_jac, sizeof(jacobian), cudaMemcpyHostToDevice, g.stream), "g1_add: H2D");
    int rc = g.msm->g1_add(d, d + 1, d + 2, g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "g1_add");
    CUDA_OR_FAIL(cudaMemcpyAsync(out_jac, d + 2, sizeof(jacobian), cudaMemcpyDeviceToHost, g.stream), "g1_add: D2H");
    CUDA_OR_FAIL(g.small_fence.leave(g.stream), "g1_add: stream order");
    CUDA_OR_FAIL(cudaStreamSynchronize(g.stream), "g1_add: execution");
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_g1_sum_device(const void* d_parts_jac, size_t count, size_t stride, size_t k, void* d_out_jac, void* stream) {
    if (!d_parts_jac || !d_out_jac) return fail(UZKGE_ERR_ARG, "g1_sum_device: null pointer");
    if (count == 0 || count > 1024 || k == 0 || k > (1u << 20) || stride < k) return fail(UZKGE_ERR_SIZE, "g1_sum_device: 1 <= count <= 1024, 1 <= k <= stride");
    API_ENTER(-1);
    return engine_fail(g.msm->g1_sum((const jacobian*)d_parts_jac, (uint32_t)count, (uint32_t)stride, (uint32_t)k, (jacobian*)d_out_jac, (cudaStream_t)stream),
                       "g1_sum_device");
}

UZKGE_API int32_t uzkge_cuda_g1_to_affine(const uint64_t in_jac[12], uint64_t out_affine[8]) {
    if (!in_jac || !out_affine) return fail(UZKGE_ERR_ARG, "g1_to_affine: null pointer");
    API_ENTER(-1);
    CUDA_OR_FAIL(g.small.reserve(4096), "g1_to_affine: buffer");
    jacobian* d = (jacobian*)g.small.p;
    CUDA_OR_FAIL(g.small_fence.enter(g.stream), "g1_to_affine: stream order");
    CUDA_OR_FAIL(cudaMemcpyAsync(d, in_jac, sizeof(jacobian), cudaMemcpyHostToDevice, g.stream), "g1_to_affine: H2D");
    int rc = g.msm->g1_to_affine(d, (affine*)(d + 1), g.stream);
    if (rc != UZKGE_OK) return engine_fail(rc, "g1_to_affine");
    CUDA_OR_FAIL(cudaMemcpyAsync(out_affine, d + 1, sizeof(affine), cudaMemcpyDeviceToHost, g.stream), "g1_to_affine: D2H");
    CUDA_OR_FAIL(g.small_fence.leave(g.stream), "g1_to_affine: stream order");
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
    API_ENTER_HANDLE(handle, "msm_g1_small_device");
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
    Profiler& prof = g_profs[g.device];
    prof.collect(sums, runs);  // drop stale records
    prof.enabled = on != 0;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_profile_read(int32_t kind, double phase_ms[8], uint64_t* runs_out) {
    if (!phase_ms || !runs_out) return fail(UZKGE_ERR_ARG, "profile_read: null pointer");
    if (kind < 0 || kind >= Profiler::KINDS) return fail(UZKGE_ERR_ARG, "profile_read: kind must be 0 (MSM) or 1 (NTT)");
    API_ENTER(-1);
    CUDA_OR_FAIL(cudaDeviceSynchronize(), "profile_read");
    static double sums[UZ_MAX_DEVICES][Profiler::KINDS][Profiler::MAX_PHASES];
    static uint64_t runs[UZ_MAX_DEVICES][Profiler::KINDS];
    g_profs[g.device].collect(sums[g.device], runs[g.device]);
    for (int i = 0; i < Profiler::MAX_PHASES; i++) {
        phase_ms[i] = sums[g.device][kind][i];
        sums[g.device][kind][i] = 0;
    }
    *runs_out = runs[g.device][kind];
    runs[g.device][kind] = 0;
    return UZKGE_OK;
}

UZKGE_API uint64_t uzkge_cuda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

UZKGE_API int32_t uzkge_cuda_configure(const char* key, uint64_t value) {
    if (!key) return fail(UZKGE_ERR_ARG, "configure: null key");
    const std::string k(key);
    if (k == "quotient_min_blocks") {
        g_quotient_min_blocks = (int)value;
        return UZKGE_OK;
    }
    if (k == "group_deal_min_log_n") {   // circuits of at least 2^value gates: a device group deals interpolations / r to its members
        if (value > 28) return fail(UZKGE_ERR_ARG, "configure: group_deal_min_log_n <= 28");
        g_group_deal_min_log_n = (int)value;
        return UZKGE_OK;
    }
    if (k == "l2_fetch_granularity") {   // bytes fetched from HBM on an L2 miss (32, 64 or 128): the MSM's 64-byte table gathers
        if (value != 32 && value != 64 && value != 128) return fail(UZKGE_ERR_ARG, "configure: l2_fetch_granularity is 32, 64 or 128");
        API_ENTER(-1);
        CUDA_OR_FAIL(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value), "configure: cudaLimitMaxL2FetchGranularity");
        return UZKGE_OK;
    }
    if (k == "virtual_devices") {   // tests: the next uzkge_cuda_init_devices builds a group of `value` members over the visible GPUs
        if (value > UZ_MAX_DEVICES) return fail(UZKGE_ERR_ARG, "configure: virtual_devices <= 16");
        g_virtual_devices = (int)value;
        return UZKGE_OK;
    }
    bool rebuild_ntt = false;
    {
        std::lock_guard<std::mutex> reg(g_reg_mu);
        if (k == "msm_lanes") {
            if (value > 32 || (value & (value - 1))) return fail(UZKGE_ERR_ARG, "configure: msm_lanes must be 0 or a power of two <= 32");
            g_cfg.msm_lanes = (uint32_t)value;
        } else if (k == "ntt_radix4") {
            if (value > 28) return fail(UZKGE_ERR_ARG, "configure: ntt_radix4 is a minimum log2 size (0 = off), at most 28");
            g_cfg.ntt_radix4 = (uint32_t)value;
        } else if (k == "ntt_big_threads") {
            if (value != 512 && value != 1024) return fail(UZKGE_ERR_ARG, "configure: ntt_big_threads is 512 or 1024");
            g_cfg.ntt_big_threads = (uint32_t)value;
        } else if (k == "ntt_log_tile") {
            if (value < 4 || value > 12) return fail(UZKGE_ERR_ARG, "configure: ntt_log_tile in 4..12");
            g_cfg.ntt_log_tile = (uint32_t)value;
            rebuild_ntt = true;
        } else if (k == "ntt_max_log_r") {
            if (value < 4 || value > 12) return fail(UZKGE_ERR_ARG, "configure: ntt_max_log_r in 4..12");
            g_cfg.ntt_max_log_r = (uint32_t)value;
            rebuild_ntt = true;
        } else if (k == "ntt_two_pass_max") {
            g_cfg.ntt_two_pass_max = (uint32_t)value;
            rebuild_ntt = true;
        } else if (k == "msm_affine") {
            g_cfg.msm_affine = (uint32_t)value;
        } else {
            return fail(UZKGE_ERR_ARG, "configure: unknown key");
        }
    }
    // apply to every device that is already up (devices initialised later read g_cfg)
    for (int d = 0; d < UZ_MAX_DEVICES; d++) {
        State* st;
        {
            std::lock_guard<std::mutex> reg(g_reg_mu);
            st = g_states[d];
        }
        if (!st) continue;
        std::lock_guard<std::mutex> lock(st->mu);
        cudaSetDevice(st->device);
        st->msm->force_lanes(g_cfg.msm_lanes);
        st->msm->set_affine(g_cfg.msm_affine);
        if (rebuild_ntt) {  // plans are cached per size: start over with the new limits
            cudaDeviceSynchronize();
            st->ntt.reset(new NttEngine(st->sm_count));
            st->ntt->configure(g_cfg.ntt_log_tile, g_cfg.ntt_max_log_r, g_cfg.ntt_two_pass_max);
        }
        st->ntt->set_big_threads(g_cfg.ntt_big_threads);
        st->ntt->set_radix4(g_cfg.ntt_radix4);
    }
    return UZKGE_OK;
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

// ---- several GPUs behind one process (`uzkge_cuda_init(device_count)`; 8e rows 1 and 2) ---------------------------------
// The reference is ONE process (shuffle/src/sdk.rs:196-227 serialises proofs behind a mutex), so a Rust host binding this library
// reaches the other GPUs of the box only if the library drives them itself.  A multi-device SRS handle owns one ordinary handle per
// device; every call on it fans out to one worker thread per device (launch overhead and the H2D copies of the devices run side by
// side) and the 96-byte partial results are combined on the host with G - 1 projective additions (ec.cuh's host twins).
namespace {

struct Pool {   // one persistent worker per device; leaked on purpose (threads blocked in a condition variable at exit)
    struct Slot {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::function<void()> fn;
        bool has = false, done = true;
    };
    std::vector<Slot*> slots;
    void ensure(size_t n) {
        while (slots.size() < n) {
            Slot* sl = new Slot();
            sl->th = std::thread([sl] {
                for (;;) {
                    std::function<void()> fn;
                    {
                        std::unique_lock<std::mutex> lk(sl->mu);
                        sl->cv.wait(lk, [sl] { return sl->has; });
                        fn = std::move(sl->fn);
                        sl->has = false;
                    }
                    fn();
                    {
                        std::lock_guard<std::mutex> lk(sl->mu);
                        sl->done = true;
                    }
                    sl->cv.notify_all();
                }
            });
            sl->th.detach();
            slots.push_back(sl);
        }
    }
    void start(size_t i, std::function<void()> fn) {
        Slot* sl = slots[i];
        {
            std::lock_guard<std::mutex> lk(sl->mu);
            sl->fn = std::move(fn);
            sl->has = true;
            sl->done = false;
        }
        sl->cv.notify_all();
    }
    void wait(size_t i) {
        Slot* sl = slots[i];
        std::unique_lock<std::mutex> lk(sl->mu);
        sl->cv.wait(lk, [sl] { return sl->done; });
    }
};

struct MultiSrs {
    int mode = 0;                      // UZKGE_MULTI_SPLIT / UZKGE_MULTI_REPLICATED
    size_t n = 0;
    std::vector<int> devices;
    std::vector<uint64_t> sub;         // the device's own handle
    std::vector<size_t> lo, hi;        // split: the device holds bases [lo, hi)
};
std::mutex g_multi_mu;                 // one multi-device call at a time (they share the workers)
Pool* g_pool = nullptr;
std::vector<int> g_group;              // devices of uzkge_cuda_init_devices
std::map<uint64_t, MultiSrs> g_multi;
uint64_t g_multi_next = 1;

struct JobResult {
    int rc = UZKGE_OK;
    std::string err;
};
// run job(i) for i < count on the workers, wait for all; the first failure (with its thread's message) is reported
int fan_out(size_t count, const std::function<int(size_t)>& job) {
    if (!g_pool) g_pool = new Pool();
    g_pool->ensure(count);
    std::vector<JobResult> res(count);
    for (size_t i = 0; i < count; i++)
        g_pool->start(i, [i, &res, &job] {
            res[i].rc = job(i);
            if (res[i].rc != UZKGE_OK) res[i].err = t_error;
        });
    for (size_t i = 0; i < count; i++) g_pool->wait(i);
    for (size_t i = 0; i < count; i++)
        if (res[i].rc != UZKGE_OK) {
            t_error = res[i].err;
            return res[i].rc;
        }
    return UZKGE_OK;
}

xyzz host_jac_to_xyzz(const uint64_t* j) {
    jacobian p;
    memcpy(&p, j, sizeof(jacobian));
    if (fe_is_zero(p.z)) return xyzz_identity();
    xyzz r;
    r.x = p.x;
    r.y = p.y;
    r.zz = fe_sqr<FqP>(p.z);
    r.zzz = fe_mul<FqP>(r.zz, p.z);
    return r;
}
// out = sum of `count` Jacobian points (12 words each): the G - 1 projective additions that merge per-GPU partial sums
void host_sum_jacobians(const uint64_t* parts, size_t count, uint64_t out[12]) {
    xyzz acc = xyzz_identity();
    for (size_t i = 0; i < count; i++) {
        const xyzz q = host_jac_to_xyzz(parts + 12 * i);
        xyzz_add(acc, q);
    }
    const jacobian r = xyzz_to_jacobian(acc);
    memcpy(out, &r, sizeof(jacobian));
}

MultiSrs* find_multi(uint64_t handle) {
    auto it = g_multi.find(handle);
    return it == g_multi.end() ? nullptr : &it->second;
}

}  // namespace

// ---- what prover.cu needs of the device group (declared in internal.h)
namespace uz {
bool SpinBarrier::wait() {
    const uint32_t gen0 = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == members) {
        count.store(0, std::memory_order_relaxed);
        gen.fetch_add(1, std::memory_order_release);
    } else {
        while (gen.load(std::memory_order_acquire) == gen0) {
            if (aborted.load(std::memory_order_relaxed)) return false;
            std::this_thread::yield();
        }
    }
    return !aborted.load(std::memory_order_relaxed);
}
std::mutex& group_mutex() { return g_multi_mu; }
std::vector<int> group_devices_locked() { return g_group; }
int group_fan_out(size_t count, const std::function<int(size_t)>& job) { return fan_out(count, job); }
void group_sum_jacobians(const uint64_t* parts, size_t count, uint64_t out[12]) { host_sum_jacobians(parts, count, out); }
bool group_srs_parts_locked(uint64_t handle, GroupSrsParts* out) {
    MultiSrs* m = find_multi(handle);
    if (!m) return false;
    out->mode = m->mode;
    out->n = m->n;
    out->devices = m->devices;
    out->sub = m->sub;
    out->lo = m->lo;
    out->hi = m->hi;
    return true;
}
}  // namespace uz

extern "C" {

UZKGE_API int32_t uzkge_cuda_init_devices(int32_t device_count) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(UZKGE_ERR_NO_DEVICE, "no CUDA device: uzkge-b200 has no CPU path");
    }
    if (device_count < 0 || device_count > count || device_count > UZ_MAX_DEVICES)
        return fail(UZKGE_ERR_NO_DEVICE, "init_devices: more devices requested than visible");
    const int use = device_count == 0 ? (count < UZ_MAX_DEVICES ? count : UZ_MAX_DEVICES) : device_count;
    std::lock_guard<std::mutex> multi(g_multi_mu);
    for (int d = 0; d < use; d++) {
        API_ENTER(d);
        // peer access: lets one device's kernels and copies reach another's memory over NVLink (ignored where unsupported)
        for (int o = 0; o < use; o++) {
            if (o == d) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d, o) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(o, 0);
                if (e != cudaSuccess) cudaGetLastError();   // already enabled
            }
        }
    }
    g_group.clear();
    if (g_virtual_devices > 0) {
        // a group with more members than GPUs: member i lives on device i % use.  Every member still owns its SRS slice, parameter
        // copy and streams, so the group code paths run unchanged -- how the one-GPU test box exercises them
        for (int i = 0; i < g_virtual_devices; i++) g_group.push_back(i % use);
    } else {
        for (int d = 0; d < use; d++) g_group.push_back(d);
    }
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_group_size(void) {
    std::lock_guard<std::mutex> multi(g_multi_mu);
    return (int32_t)g_group.size();
}

UZKGE_API int32_t uzkge_cuda_srs_upload(const uint64_t* affine_xy, size_t n, uint32_t window_bits, uint64_t* handle) {
    return srs_upload_one(-1, affine_xy, n, window_bits, handle);
}

UZKGE_API int32_t uzkge_cuda_srs_upload_multi(const uint64_t* affine_xy, size_t n, uint32_t window_bits, int32_t mode, uint64_t* handle) {
    if (!affine_xy || !handle) return fail(UZKGE_ERR_ARG, "srs_upload_multi: null pointer");
    if (mode != UZKGE_MULTI_SPLIT && mode != UZKGE_MULTI_REPLICATED) return fail(UZKGE_ERR_ARG, "srs_upload_multi: mode");
    std::lock_guard<std::mutex> multi(g_multi_mu);
    if (g_group.empty()) return fail(UZKGE_ERR_NO_DEVICE, "srs_upload_multi: call uzkge_cuda_init_devices first");
    if (n == 0) return fail(UZKGE_ERR_SIZE, "srs_upload_multi: empty SRS");
    MultiSrs m;
    m.mode = mode;
    m.n = n;
    size_t G = g_group.size();
    if (mode == UZKGE_MULTI_SPLIT && G > n) G = n;
    const size_t base = n / G, rem = n % G;
    for (size_t i = 0; i < G; i++) {
        m.devices.push_back(g_group[i]);
        const size_t lo = mode == UZKGE_MULTI_SPLIT ? i * base + (i < rem ? i : rem) : 0;
        const size_t hi = mode == UZKGE_MULTI_SPLIT ? lo + base + (i < rem ? 1 : 0) : n;
        m.lo.push_back(lo);
        m.hi.push_back(hi);
    }
    m.sub.assign(G, 0);
    int rc = fan_out(G, [&](size_t i) { return srs_upload_one(m.devices[i], affine_xy + 8 * m.lo[i], m.hi[i] - m.lo[i], window_bits, &m.sub[i]); });
    if (rc != UZKGE_OK) {
        const std::string keep = t_error;
        for (size_t i = 0; i < G; i++)
            if (m.sub[i]) srs_free_one(m.sub[i]);
        t_error = keep;
        return rc;
    }
    const uint64_t h = HANDLE_MULTI | g_multi_next++;
    g_multi[h] = m;
    *handle = h;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_srs_free(uint64_t handle) {
    if (!(handle & HANDLE_MULTI)) return srs_free_one(handle);
    std::lock_guard<std::mutex> multi(g_multi_mu);
    MultiSrs* m = find_multi(handle);
    if (!m) return fail(UZKGE_ERR_HANDLE, "srs_free: unknown handle");
    int rc = UZKGE_OK;
    for (uint64_t h : m->sub) {
        const int r = srs_free_one(h);
        if (r != UZKGE_OK) rc = r;
    }
    g_multi.erase(handle);
    return rc;
}

UZKGE_API int32_t uzkge_cuda_srs_info(uint64_t handle, uzkge_srs_info* info) {
    if (!(handle & HANDLE_MULTI)) return srs_info_one(handle, info);
    if (!info) return fail(UZKGE_ERR_ARG, "srs_info: null pointer");
    std::lock_guard<std::mutex> multi(g_multi_mu);
    MultiSrs* m = find_multi(handle);
    if (!m) return fail(UZKGE_ERR_HANDLE, "srs_info: unknown handle");
    uzkge_srs_info sum = {};
    for (size_t i = 0; i < m->sub.size(); i++) {
        uzkge_srs_info one;
        const int rc = srs_info_one(m->sub[i], &one);
        if (rc != UZKGE_OK) return rc;
        if (i == 0) sum = one;
        else {
            sum.device_bytes += one.device_bytes;
            if (one.precompute_ms > sum.precompute_ms) sum.precompute_ms = one.precompute_ms;
        }
    }
    sum.n = m->n;
    sum.reserved = (uint32_t)m->sub.size();   // number of devices behind the handle
    *info = sum;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_msm_g1(uint64_t handle, size_t base_offset, const uint64_t* scalars, size_t n, uint64_t out_jac[12]) {
    if (!(handle & HANDLE_MULTI)) return msm_g1_one(handle, base_offset, scalars, n, out_jac);
    if (!out_jac || (n && !scalars)) return fail(UZKGE_ERR_ARG, "msm_g1: null pointer");
    std::lock_guard<std::mutex> multi(g_multi_mu);
    MultiSrs* m = find_multi(handle);
    if (!m) return fail(UZKGE_ERR_HANDLE, "msm_g1: unknown handle");
    if (base_offset > m->n || n > m->n - base_offset) return fail(UZKGE_ERR_SIZE, "msm_g1: range outside the SRS");
    if (m->mode == UZKGE_MULTI_REPLICATED) return msm_g1_one(m->sub[0], base_offset, scalars, n, out_jac);
    // split: device i computes the part of [base_offset, base_offset + n) that falls into its slice
    const size_t G = m->sub.size();
    std::vector<uint64_t> parts(12 * G, 0);
    int rc = fan_out(G, [&](size_t i) {
        const size_t a = base_offset > m->lo[i] ? base_offset : m->lo[i];
        const size_t b = base_offset + n < m->hi[i] ? base_offset + n : m->hi[i];
        if (b <= a) return (int)UZKGE_OK;            // Z = 0: identity
        return (int)msm_g1_one(m->sub[i], a - m->lo[i], scalars + 4 * (a - base_offset), b - a, &parts[12 * i]);
    });
    if (rc != UZKGE_OK) return rc;
    host_sum_jacobians(parts.data(), G, out_jac);
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_msm_g1_batch(uint64_t handle, const uint64_t* const* scalars, const size_t* n, size_t k, uint64_t* out_jac) {
    if (!(handle & HANDLE_MULTI)) return msm_g1_batch_one(handle, scalars, n, k, out_jac);
    if (k && (!scalars || !n || !out_jac)) return fail(UZKGE_ERR_ARG, "msm_g1_batch: null pointer");
    std::lock_guard<std::mutex> multi(g_multi_mu);
    MultiSrs* m = find_multi(handle);
    if (!m) return fail(UZKGE_ERR_HANDLE, "msm_g1: unknown handle");
    if (k == 0) return UZKGE_OK;
    for (size_t j = 0; j < k; j++) {
        if (n[j] > m->n) return fail(UZKGE_ERR_SIZE, "msm_g1: more scalars than SRS points");
        if (n[j] && !scalars[j]) return fail(UZKGE_ERR_ARG, "msm_g1: null scalar vector");
    }
    const size_t G = m->sub.size();
    if (m->mode == UZKGE_MULTI_REPLICATED) {
        // the independent commitments of a round (plonk/prover.rs:132-192, helpers.rs:1323-1408) dealt to the devices: MSM j on
        // device j % G, each device runs its share as one batch
        std::vector<std::vector<const uint64_t*>> sc(G);
        std::vector<std::vector<size_t>> nn(G), idx(G);
        std::vector<std::vector<uint64_t>> outs(G);
        for (size_t j = 0; j < k; j++) {
            sc[j % G].push_back(scalars[j]);
            nn[j % G].push_back(n[j]);
            idx[j % G].push_back(j);
        }
        for (size_t i = 0; i < G; i++) outs[i].assign(12 * sc[i].size() + 12, 0);
        int rc = fan_out(G, [&](size_t i) {
            if (sc[i].empty()) return (int)UZKGE_OK;
            return (int)msm_g1_batch_one(m->sub[i], sc[i].data(), nn[i].data(), sc[i].size(), outs[i].data());
        });
        if (rc != UZKGE_OK) return rc;
        for (size_t i = 0; i < G; i++)
            for (size_t t = 0; t < idx[i].size(); t++) memcpy(out_jac + 12 * idx[i][t], &outs[i][12 * t], 96);
        return UZKGE_OK;
    }
    // split: every device runs all k MSMs over its slice (a prefix of the SRS meets a slice in a prefix of the slice)
    std::vector<std::vector<uint64_t>> parts(G, std::vector<uint64_t>(12 * k, 0));
    int rc = fan_out(G, [&](size_t i) {
        std::vector<const uint64_t*> sc(k);
        std::vector<size_t> nn(k);
        for (size_t j = 0; j < k; j++) {
            const size_t b = n[j] < m->hi[i] ? n[j] : m->hi[i];
            nn[j] = b > m->lo[i] ? b - m->lo[i] : 0;
            sc[j] = nn[j] ? scalars[j] + 4 * m->lo[i] : nullptr;
        }
        return (int)msm_g1_batch_one(m->sub[i], sc.data(), nn.data(), k, parts[i].data());
    });
    if (rc != UZKGE_OK) return rc;
    std::vector<uint64_t> col(12 * G);
    for (size_t j = 0; j < k; j++) {
        for (size_t i = 0; i < G; i++) memcpy(&col[12 * i], &parts[i][12 * j], 96);
        host_sum_jacobians(col.data(), G, out_jac + 12 * j);
    }
    return UZKGE_OK;
}

// One transform over the whole device group, host pointers in and out ("NTTs of size 2^22 and above use a four-step transpose over
// NVLink"): member r uploads slice r of the input over ITS host link, the cross-rank kernel loads its column block straight from every
// member's slice and stores row k1 into member k1's buffer (peer memory: both exchanges are the kernel's own loads and stores), the
// size-n/G local transform's last pass stores into the owners' natural slices, and member r returns slice r of the result.  Three
// host-side barriers order the members.  Falls back to the single-device call where the four-step does not apply (group of 1, 3, 5..
// members, 3 * 2^k domains, sizes below G^2).
namespace {
struct GroupNttBufs {
    DevBuf x, rows, nat, tmp;
};
std::vector<GroupNttBufs> g_group_ntt;      // one per group member (guarded by g_multi_mu)
}  // namespace

UZKGE_API int32_t uzkge_cuda_ntt_fr_multi(uint64_t* inout, size_t len_in, size_t domain_size, int32_t inverse, const uint64_t* coset_shift) {
    if (!inout) return fail(UZKGE_ERR_ARG, "ntt_fr_multi: null pointer");
    if (inverse != 0 && inverse != 1) return fail(UZKGE_ERR_ARG, "ntt_fr_multi: inverse must be 0 or 1");
    if (len_in > domain_size) return fail(UZKGE_ERR_SIZE, "ntt_fr_multi: input longer than the domain");
    std::lock_guard<std::mutex> multi(g_multi_mu);
    const size_t G = g_group.size();
    const size_t n = domain_size;
    uint32_t log_g = 0;
    while ((1ull << log_g) < G) log_g++;
    if (G < 2 || G > 8 || (1ull << log_g) != G || n == 0 || (n & (n - 1)) || n < G * G)
        return uzkge_cuda_ntt_fr(inout, len_in, domain_size, inverse, coset_shift);
    const size_t L = n / G, S = L / G;
    if (g_group_ntt.size() != G) {     // the group changed: the old members' buffers go (a DevBuf does not free itself)
        for (GroupNttBufs& b : g_group_ntt)
            for (DevBuf* d : {&b.x, &b.rows, &b.nat, &b.tmp})
                if (d->p) {
                    cudaPointerAttributes at;
                    if (cudaPointerGetAttributes(&at, d->p) == cudaSuccess && cudaSetDevice(at.device) == cudaSuccess) cudaFree(d->p);
                    cudaGetLastError();
                }
        g_group_ntt.assign(G, GroupNttBufs());
    }
    fe shift = fe_one<FrP>();
    if (coset_shift) memcpy(&shift, coset_shift, sizeof(fe));
    std::vector<fe*> X(G, nullptr), R(G, nullptr), N(G, nullptr);
    SpinBarrier bar;
    bar.members = (uint32_t)G;
    // one phase of member r under its device's lock (never held across a barrier: two members may share a device in a virtual group)
    auto phase = [&](size_t r, const std::function<int(State&, GroupNttBufs&)>& body) -> int {
        State* st = nullptr;
        int rc = enter_state(g_group[r], &st);
        if (rc != UZKGE_OK) return rc;
        std::lock_guard<std::mutex> lock(st->mu);
        cudaError_t e = cudaSetDevice(st->device);
        if (e != cudaSuccess) return fail_cuda("cudaSetDevice", e);
        return body(*st, g_group_ntt[r]);
    };
    // x[j] *= shift^(first + j) on `count` elements (the coset scaling of mul_var_assign, field_polynomial.rs:470-477)
    auto scale_by_powers = [&](State& dv, fe* v, fe* tmp, size_t first, size_t count) -> int {
        const fe start = fe_pow_u64<FrP>(shift, first);
        int rc = fr_powers_run((const uint64_t*)&shift, (const uint64_t*)&start, count, tmp, dv.stream);
        if (rc != UZKGE_OK) return rc;
        return fr_mul_run(v, tmp, count, v, dv.stream);
    };
    auto member = [&](size_t r) -> int {
        int rc = phase(r, [&](State& dv, GroupNttBufs& b) -> int {
            CUDA_OR_FAIL(b.x.reserve(L * sizeof(fe)), "ntt_fr_multi: buffer");
            CUDA_OR_FAIL(b.rows.reserve(L * sizeof(fe)), "ntt_fr_multi: buffer");
            CUDA_OR_FAIL(b.nat.reserve(L * sizeof(fe)), "ntt_fr_multi: buffer");
            CUDA_OR_FAIL(b.tmp.reserve(L * sizeof(fe)), "ntt_fr_multi: buffer");
            X[r] = (fe*)b.x.p;
            R[r] = (fe*)b.rows.p;
            N[r] = (fe*)b.nat.p;
            const size_t lo = r * L;
            const size_t have = len_in > lo ? (len_in - lo < L ? len_in - lo : L) : 0;
            if (have) CUDA_OR_FAIL(cudaMemcpyAsync(X[r], inout + 4 * lo, have * sizeof(fe), cudaMemcpyHostToDevice, dv.stream), "ntt_fr_multi: H2D");
            if (have < L) CUDA_OR_FAIL(cudaMemsetAsync(X[r] + have, 0, (L - have) * sizeof(fe), dv.stream), "ntt_fr_multi: zero padding");
            if (coset_shift && !inverse && have) {
                const int rc2 = scale_by_powers(dv, X[r], (fe*)b.tmp.p, lo, have);
                if (rc2 != UZKGE_OK) return engine_fail(rc2, "ntt_fr_multi: coset scaling");
            }
            CUDA_OR_FAIL(cudaStreamSynchronize(dv.stream), "ntt_fr_multi: upload");
            return UZKGE_OK;
        });
        if (rc != UZKGE_OK) return rc;
        if (!bar.wait()) return fail(UZKGE_ERR_INTERNAL, "ntt_fr_multi: another member of the device group failed");
        rc = phase(r, [&](State& dv, GroupNttBufs&) -> int {
            const fe* in_rows[8];
            fe* out_rows[8];
            for (size_t i = 0; i < G; i++) {
                in_rows[i] = X[i] + r * S;
                out_rows[i] = R[i] + r * S;
            }
            const int rc2 = dv.ntt->cross_rows(in_rows, out_rows, log_g, S, r * S, n, inverse != 0, dv.stream);
            if (rc2 != UZKGE_OK) return engine_fail(rc2, "ntt_fr_multi: cross step");
            CUDA_OR_FAIL(cudaStreamSynchronize(dv.stream), "ntt_fr_multi: cross step");
            return UZKGE_OK;
        });
        if (rc != UZKGE_OK) return rc;
        if (!bar.wait()) return fail(UZKGE_ERR_INTERNAL, "ntt_fr_multi: another member of the device group failed");
        rc = phase(r, [&](State& dv, GroupNttBufs& b) -> int {
            NttScatter sc;
            for (size_t i = 0; i < 8; i++) sc.rows[i] = i < G ? N[i] : nullptr;
            sc.log_g = log_g;
            sc.k1 = (uint32_t)r;
            const fe* in = R[r];
            fe* out = (fe*)b.tmp.p;
            const uint64_t len = L;
            const int rc2 = dv.ntt->run_batch(&in, &out, (fe*)b.tmp.p, &len, 1, L, inverse != 0, nullptr, dv.stream, &sc);
            if (rc2 != UZKGE_OK) return engine_fail(rc2, "ntt_fr_multi: local transform");
            CUDA_OR_FAIL(cudaStreamSynchronize(dv.stream), "ntt_fr_multi: local transform");
            return UZKGE_OK;
        });
        if (rc != UZKGE_OK) return rc;
        if (!bar.wait()) return fail(UZKGE_ERR_INTERNAL, "ntt_fr_multi: another member of the device group failed");
        return phase(r, [&](State& dv, GroupNttBufs& b) -> int {
            if (coset_shift && inverse) {
                const int rc2 = scale_by_powers(dv, N[r], (fe*)b.tmp.p, r * L, L);
                if (rc2 != UZKGE_OK) return engine_fail(rc2, "ntt_fr_multi: coset scaling");
            }
            CUDA_OR_FAIL(cudaMemcpyAsync(inout + 4 * r * L, N[r], L * sizeof(fe), cudaMemcpyDeviceToHost, dv.stream), "ntt_fr_multi: D2H");
            CUDA_OR_FAIL(cudaStreamSynchronize(dv.stream), "ntt_fr_multi: execution");
            return UZKGE_OK;
        });
    };
    return fan_out(G, [&](size_t r) {
        const int rc = member(r);
        if (rc != UZKGE_OK) bar.abort();
        return rc;
    });
}

}  // extern "C"
