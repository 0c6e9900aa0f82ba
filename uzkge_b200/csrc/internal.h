// internal.h -- host-side engine declarations shared by ntt.cu, msm.cu and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <functional>
#include <mutex>
#include <map>
#include <vector>

#include "../../include/uzkge_cuda.h"
#include "ec.cuh"
#include "ntt_plan.h"

namespace uz {

extern std::atomic<uint64_t> g_launches;  // kernels launched by this library (uzkge_cuda_launch_count)
#define UZ_COUNT_LAUNCH(k) (::uz::g_launches.fetch_add((k), std::memory_order_relaxed))

// ---------------------------------------------------------------- per-phase device timing (bench / profiling)
// When enabled, every engine run records CUDA events on its launching stream at phase boundaries; the sums
// are read back (after a synchronise) through uzkge_cuda_profile_read.  Off by default: no events, no cost.
struct Profiler {
    static constexpr int MAX_PHASES = 8;
    enum Kind { MSM = 0, NTT = 1, KINDS = 2 };
    struct Rec {
        int kind;
        int nmarks;
        int phase_of_mark[MAX_PHASES + 1];
        cudaEvent_t ev[MAX_PHASES + 1];
    };
    bool enabled = false;
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    // returns a record index (or -1 when disabled) and records the start event
    int begin(int kind, cudaStream_t st) {
        if (!enabled || recs.size() >= 65536) return -1;
        Rec r;
        r.kind = kind;
        r.nmarks = 1;
        r.phase_of_mark[0] = -1;
        r.ev[0] = get();
        cudaEventRecord(r.ev[0], st);
        recs.push_back(r);
        return (int)recs.size() - 1;
    }
    // everything launched since the previous mark belongs to `phase`
    void mark(int rec, int phase, cudaStream_t st) {
        if (rec < 0) return;
        Rec& r = recs[rec];
        if (r.nmarks > MAX_PHASES) return;
        r.ev[r.nmarks] = get();
        r.phase_of_mark[r.nmarks] = phase;
        cudaEventRecord(r.ev[r.nmarks], st);
        r.nmarks++;
    }
    // sums[kind][phase] += elapsed; runs[kind]++; releases the events.  Caller has synchronised the device.
    void collect(double sums[KINDS][MAX_PHASES], uint64_t runs[KINDS]) {
        for (Rec& r : recs) {
            for (int i = 1; i < r.nmarks; i++) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, r.ev[i - 1], r.ev[i]) == cudaSuccess) sums[r.kind][r.phase_of_mark[i]] += ms;
            }
            for (int i = 0; i < r.nmarks; i++) pool.push_back(r.ev[i]);
            runs[r.kind]++;
        }
        recs.clear();
    }
};
// one profiler per device (events belong to the device they were created on); engines reach theirs through the CUDA runtime's
// current device, which every entry point has set before it calls them
static constexpr int UZ_MAX_DEVICES = 16;
Profiler& prof_for_current_device();
#define g_prof (::uz::prof_for_current_device())
enum MsmPhase { MSM_PH_RECODE = 0, MSM_PH_SORT, MSM_PH_OFFSETS, MSM_PH_ACCUMULATE, MSM_PH_LARGE, MSM_PH_REDUCE };
enum NttPhase { NTT_PH_RADIX3 = 0, NTT_PH_PASS0, NTT_PH_PASS1, NTT_PH_PASS2 };

// ---------------------------------------------------------------- NTT
struct NttDomain {
    NttPlan plan;
    fe omega, n_inv, w3, w3sq;
    fe* w_lo = nullptr;   // one allocation: w_lo | w_hi | stage
    fe* w_hi = nullptr;
    fe* stage = nullptr;
    uint32_t n_hi = 0, n_stage = 0;
    fe* tw_all = nullptr;                          // one allocation: the per-element twiddle tables below
    fe* tw_pass[3] = {nullptr, nullptr, nullptr};  // inter-pass twiddles, indexed by the pass's output position
    fe* tw3 = nullptr;                             // radix-3 pre-pass twiddles w^n | w^(2n)
};
struct NttCoset {
    uint64_t n;
    fe g;
    bool with_ninv;
    fe* g_lo = nullptr;   // one allocation: g_lo | g_hi
    fe* g_hi = nullptr;
    fe* g_full = nullptr;  // premul * g^j for every j < n
    uint32_t n_hi = 0;
};


static constexpr uint32_t NTT_MAX_BATCH = 16;

// the last pass of a distributed transform's local step stores into the owners' natural slices (peer memory): see ntt.cu
struct NttScatter {
    fe* rows[8];        // rank r's natural output slice (n elements each)
    uint32_t log_g;     // log2 of the number of ranks
    uint32_t k1;        // this rank
};

class NttEngine {
public:
    explicit NttEngine(int sm_count) : sm_count_(sm_count) {}
    ~NttEngine();
    void set_big_threads(uint32_t nt) { cfg_big_threads_ = nt; }
    void set_radix4(uint32_t on) { cfg_radix4_ = on; }
    void configure(uint32_t log_tile, uint32_t max_log_r, uint32_t two_pass_max) {
        cfg_log_tile_ = log_tile;
        cfg_max_log_r_ = max_log_r;
        cfg_two_pass_max_ = two_pass_max;
    }
    int run(const fe* d_in, fe* d_out, fe* d_scratch, uint64_t len_in, uint64_t n, bool inverse, const fe* coset_shift,
            cudaStream_t st);
    // k <= NTT_MAX_BATCH vectors over one domain, one launch per pass (blockIdx.y = vector); d_scratch: k * n elements
    int run_batch(const fe* const* d_in, fe* const* d_out, fe* d_scratch, const uint64_t* len_in, uint32_t k, uint64_t n, bool inverse,
                  const fe* coset_shift, cudaStream_t st, const NttScatter* scatter = nullptr);
    // cross-rank step of a distributed transform (see ntt.cu)
    int cross(const fe* d_in, fe* d_out, uint32_t log_g, uint64_t cols, uint64_t col_offset, uint64_t n_total, bool inverse,
              cudaStream_t st);
    int cross_rows(const fe* const* in_rows, fe* const* out_rows, uint32_t log_g, uint64_t cols, uint64_t col_offset, uint64_t n_total,
                   bool inverse, cudaStream_t st);

private:
    const NttDomain* domain(uint64_t n, cudaStream_t st);
    const NttCoset* coset(uint64_t n, const fe& g, bool with_ninv, const NttDomain* d, cudaStream_t st);
    std::map<uint64_t, NttDomain> domains_;
    std::vector<NttCoset> cosets_;
    int sm_count_;
    uint32_t cfg_log_tile_ = 10, cfg_max_log_r_ = 10, cfg_two_pass_max_ = 18, cfg_big_threads_ = 1024, cfg_radix4_ = 19;
};

// ---------------------------------------------------------------- MSM
struct ReducePlan;
// One set of per-pass buffers, sized for `slots` MSMs over the whole SRS.  An SRS owns two: a run of more MSMs than fit one pass
// alternates between them so that the counting sort of group g + 1 and the bucket reduction of group g - 1 overlap the accumulate
// kernel of group g (MsmEngine::run_pipelined).
struct MsmWork {
    uint32_t* vals = nullptr;         // entries sorted by global bucket id: sign | (f * n + point)
    uint32_t* offsets = nullptr;      // slots * nbuckets + 1
    uint32_t* counts = nullptr;       // slots * nbuckets + 1
    uint32_t *ord_keys_a = nullptr, *ord_keys_b = nullptr, *ord_vals_a = nullptr, *ord_vals_b = nullptr;  // size-ordered bucket ids
    uint32_t* large_list = nullptr;   // [0] = count, then bucket ids
    uint32_t* slice_start = nullptr;  // prefix of warp slices per oversized bucket
    xyzz* slice_sums = nullptr;
    xyzz* buckets = nullptr;          // slots * nbuckets
    ReducePlan* reduce[16] = {};      // launch plan of the bucket reduction of every slot (msm_reduce.cu)
    void* cub_temp = nullptr;
    void* aff_ws = nullptr;           // point arrays + prefix products of the batched-affine accumulation (msm_affine.cu), on demand
    size_t aff_ws_bytes = 0;
    cudaEvent_t sorted = nullptr, accumulated = nullptr, reduced = nullptr;  // pipeline hand-offs
};
struct MsmSrs {
    uint64_t n = 0;            // points
    uint32_t c = 0;            // window bits
    uint32_t windows = 0;      // number of windows == number of fixed-base tables
    uint32_t nbuckets = 0;     // 2^(c-1) + 1 bucket ids per MSM (id 0 = zero digit, always empty)
    uint32_t slots = 1;        // independent MSMs one pass can carry (batch)
    uint32_t large_cap = 0, max_slices = 0;
    double precompute_ms = 0;
    void* arena = nullptr;     // one device allocation holding everything below
    size_t bytes = 0;
    affine* tables = nullptr;  // windows x n affine points: table f holds 2^(c*f) * P_i
    MsmWork work[2];
    size_t cub_temp_bytes = 0;
    size_t reduce_bytes = 0;   // reduction workspace of one slot (the slots' workspaces are contiguous)
};

// inverse transform over G1 points: Lagrange SRS from the monomial SRS (ec_ntt.cu)
int ec_intt_run(const affine* d_points, uint32_t n, uint32_t log_n, const fe* d_tw, const fe& n_inv_canonical, xyzz* d_work, affine* d_out,
                cudaStream_t st);

// batched-affine bucket accumulation (msm_affine.cu)
size_t msm_affine_workspace_bytes(uint64_t m, uint32_t nb);
uint32_t msm_affine_rounds(double mean_load);
int msm_affine_accumulate(const MsmSrs* s, const MsmWork& w, void* aff_ws, uint64_t m, uint32_t nb, const uint32_t* order, uint32_t thr,
                          uint32_t rounds, cudaStream_t st, cudaStream_t st2, cudaEvent_t fork, cudaEvent_t join);

// bucket reduction (msm_reduce.cu): workspace size for window size c, plan over fixed buffers, run
size_t msm_reduce_workspace_bytes(uint32_t c, uint32_t sm_count);
ReducePlan* msm_reduce_plan_create(uint32_t c, uint32_t sm_count, xyzz* buckets, void* workspace, uint32_t* ticket);
void msm_reduce_plan_destroy(ReducePlan* p);
int msm_reduce_run(const ReducePlan* p, uint32_t k, const xyzz* buckets, size_t bucket_stride, size_t ws_stride_bytes, jacobian* d_out,
                   cudaStream_t st);

class MsmEngine {
public:
    explicit MsmEngine(int sm_count) : sm_count_(sm_count) {}
    ~MsmEngine();
    int upload(const uint64_t* affine_xy_host, size_t n, uint32_t window_bits, MsmSrs* out, cudaStream_t st);
    void release(MsmSrs* s);
    // d_scalars: n Montgomery Fr on the device; d_out: 12 x u64 Jacobian on the device
    int run(MsmSrs* s, size_t base_offset, const fe* d_scalars, size_t n, jacobian* d_out, cudaStream_t st);
    // k <= s->slots independent MSMs in one pass (one sort, one accumulate launch, concurrent reductions)
    int run_batch(MsmSrs* s, size_t base_offset, const fe* const* d_scalars, const size_t* n, uint32_t k, jacobian* d_out,
                  cudaStream_t st);
    // any number of independent MSMs: groups of s->slots, software-pipelined over the SRS's two workspaces on internal streams
    // (sort of group g + 1 and reduction of group g - 1 run under the accumulate kernel of group g); joins `st` at the end
    // `ready` (optional, k events): MSM j's scalars are complete once ready[j] has fired (host-pointer API: the H2D copy of group
    // g + 1 then runs under the kernels of group g)
    int run_pipelined(MsmSrs* s, size_t base_offset, const fe* const* d_scalars, const size_t* n, size_t k, jacobian* d_out,
                      cudaStream_t st, const cudaEvent_t* ready = nullptr);
    int g1_add(const jacobian* d_a, const jacobian* d_b, jacobian* d_out, cudaStream_t st);
    int g1_sum(const jacobian* d_parts, uint32_t count, uint32_t stride, uint32_t k, jacobian* d_out, cudaStream_t st);
    int g1_to_affine(const jacobian* d_in, affine* d_out, cudaStream_t st);
    int powers_of_tau(const fe& tau, uint64_t first, uint32_t count, affine* d_out, cudaStream_t st);
    int fixed_base_mul(const fe* d_scalars, uint32_t count, affine* d_out, cudaStream_t st);
    // *d_out (+)= sum_j scalars[j] * srs[idx[j]], k <= 32 (host scalars, Montgomery)
    int small_msm(const MsmSrs* s, const size_t* idx, const uint64_t* scalars, uint32_t k, bool accumulate, jacobian* d_out, cudaStream_t st);
    void force_lanes(uint32_t g) { force_lanes_ = g; }  // tuning knob (0 = automatic)
    void set_affine(uint32_t mode) { affine_mode_ = mode; }  // 0 = XYZZ accumulation only, 1 = batched-affine tree where it pays


private:
    struct GroupPlan;
    int stage_sort(MsmSrs* s, MsmWork& w, size_t base_offset, const fe* const* d_scalars, const size_t* n, uint32_t k, GroupPlan* plan,
                   cudaStream_t st, int prof);
    int stage_accumulate(MsmSrs* s, MsmWork& w, const GroupPlan& plan, cudaStream_t st, int prof);
    int stage_reduce(MsmSrs* s, MsmWork& w, uint32_t k, jacobian* d_out, cudaStream_t st);
    int sm_count_;
    cudaStream_t s_sort_ = nullptr, s_acc_ = nullptr, s_red_ = nullptr;  // the pipeline's streams
    cudaEvent_t start_ = nullptr;
    cudaStream_t aff_stream_ = nullptr;   // second stream of the batched-affine accumulation
    cudaEvent_t aff_fork_ = nullptr, aff_join_ = nullptr;
    uint32_t force_lanes_ = 0;
    uint32_t affine_mode_ = 0;
    std::vector<cudaStream_t> aux_;   // auxiliary streams for the concurrent reductions of a batch
    std::vector<cudaEvent_t> join_;
    cudaEvent_t fork_ = nullptr;
};

// ---------------------------------------------------------------- polynomial glue (poly.cu)
class PolyEngine {
public:
    ~PolyEngine();
    // T_k = c_k + z T_{k+1}: *d_value = T_0 = p(z);  d_quot (n - 1 coefficients, may be null) = p / (X - z)
    int horner(const fe* d_c, uint64_t n, const fe& z, fe* d_quot, fe* d_value, cudaStream_t st);
    // d_values[j] = p_j(points[point[j]]), k pairs in two launches
    int eval_batch(const fe* const* d_c, const uint64_t* n, const uint32_t* point, uint32_t k, const fe* points, uint32_t npoints,
                   fe* d_values, cudaStream_t st);
    // d_out[0] = 1, d_out[i + 1] = prod_{j <= i} num_j / den_j  (n + 1 outputs);  d_tmp: 2 n + 2 elements
    int grand_product(const fe* d_num, const fe* d_den, uint64_t n, fe* d_out, fe* d_tmp, cudaStream_t st);

private:
    void* ws_ = nullptr;
    size_t ws_cap_ = 0;
};

// ---------------------------------------------------------------- PlonK quotient map (quotient.cu)
int plonk_quotient_run(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* shuffle, void* d_out, cudaStream_t st);
// the same map on the points start + step * idx, idx < count, of the size-m arrays
int plonk_quotient_range_run(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* shuffle, uint64_t start, uint64_t step,
                             uint64_t count, void* d_out, cudaStream_t st);

// ---------------------------------------------------------------- elementwise prover glue (plonk_glue.cu)
int fr_lincomb_run(const void* const* d_polys, const size_t* lens, const uint64_t* coefs, size_t k, void* d_out, size_t out_len, cudaStream_t st);
int fr_add_sparse_run(void* d_poly, const size_t* idx, const uint64_t* vals, size_t k, cudaStream_t st);
int fr_add_sparse_multi_run(void* const* d_polys, const size_t* idx, const uint64_t* vals, size_t k, cudaStream_t st);
int fr_powers_run(const uint64_t* base, const uint64_t* scale, size_t n, void* d_out, cudaStream_t st);
int fr_gather_run(const void* d_src, const void* d_idx, size_t n, void* d_out, cudaStream_t st);
int fr_gather_scatter_run(const void* d_src, const void* d_src_idx, void* d_dst, const void* d_dst_idx, size_t k, cudaStream_t st);
int fr_mul_run(const void* d_a, const void* d_b, size_t n, void* d_out, cudaStream_t st);
// t's f n coefficients from the f per-coset inverse transforms u_j (f x n, compact); k1: the quotient coset's shift (host, Montgomery)
int plonk_coset_combine_run(const void* d_u, size_t n, size_t factor, const uint64_t* k1, void* d_out, cudaStream_t st);
// d_dst[dst_start + dst_step * i] = d_src[src_start + src_step * i], i < count
int fr_strided_copy_run(const void* d_src, size_t src_start, size_t src_step, void* d_dst, size_t dst_start, size_t dst_step, size_t count,
                        cudaStream_t st);
int fr_trimmed_len_run(const void* d_poly, size_t n, unsigned long long* d_scratch, size_t* len_out, cudaStream_t st);
int plonk_z_evals_run(PolyEngine* poly, const void* const d_w[5], const void* const d_sigma[5], const void* d_group, const uint64_t* k,
                      const uint64_t* beta, const uint64_t* gamma, size_t n, void* d_z, void* d_tmp, cudaStream_t st);

// The members of a device group are dedicated worker threads that meet a few times per call: spin, do not sleep.  wait() returns
// false when a member failed and the call is abandoned.
struct SpinBarrier {
    std::atomic<uint32_t> count{0}, gen{0};
    std::atomic<bool> aborted{false};
    uint32_t members = 1;
    bool wait();
    void abort() { aborted.store(true); }
};

// ---------------------------------------------------------------- the device group (api.cu) as prover.cu sees it
struct GroupSrsParts {      // a multi-device SRS handle: member i owns bases [lo[i], hi[i]) behind its own single-device handle
    int mode = 0;
    size_t n = 0;
    std::vector<int> devices;
    std::vector<uint64_t> sub;
    std::vector<size_t> lo, hi;
};
std::mutex& group_mutex();                       // one multi-device call at a time (the calls share the worker threads)
std::vector<int> group_devices_locked();         // members' devices (group_mutex held)
bool group_srs_parts_locked(uint64_t handle, GroupSrsParts* out);
int group_fan_out(size_t count, const std::function<int(size_t)>& job);   // job(i) on worker i, all in parallel; first failure reported
void group_sum_jacobians(const uint64_t* parts, size_t count, uint64_t out[12]);

static inline int cuda_err_code(cudaError_t e) { return e == cudaErrorMemoryAllocation ? UZKGE_ERR_OOM : UZKGE_ERR_CUDA; }

}  // namespace uz
