// prover.cu -- uzkge_cuda_plonk_*: the TurboPlonK prover of the reference behind the C ABI.
//
// What in production is the patched body of `prover_with_lagrange` (/root/reference/uzkge/src/plonk/prover.rs:88-394) calls
// uzkge_cuda_plonk_prove ONCE: the five rounds below run device-resident on the library's own kernels (through the same entry points
// a host would call one by one), and the Fiat-Shamir transcript (utils/transcript.rs:8-69), the O(1) scalar arithmetic of every
// round (helpers.rs:681-999, 1412-1423) and the order of operations are restated here in C++:
//
//   pi_poly, hide_polynomial, z_poly, t_poly, split_t_and_commit, r_poly   plonk/helpers.rs:111-131, 139-154, 160-220, 223-678, 1323-1408, 681-999
//   batch_prove                                                            poly_commit/pcs.rs:107-168
//   the `commit` closure of the Lagrange branch                            plonk/prover.rs:131-146, kzg_poly_commitment.rs:299-313
//
// The Python mirror (uzkge_b200/plonk.py::prover) is the same algorithm call for call; tests/test_gpu_prover_native.py holds the two
// to identical proof bytes, and both to the big-integer restatement of the reference (oracle/plonk_prover.py).
// Host code only: no kernel lives here.  No CPU fallback: every polynomial operation is a device call.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/uzkge_transcript.hpp"
#include "internal.h"

namespace uz {
int api_fail(int code, const char* what);   // api.cu: sets the calling thread's error message
// device groups deal a round's interpolations and the linearisation polynomial to their members from this circuit size on (below it a
// round's transforms are one batched launch per pass and the extra barriers cost more than they save); uzkge_cuda_configure
// "group_deal_min_log_n"
int g_group_deal_min_log_n = 19;
}

namespace {

using uzkge::FQ;
using uzkge::FR;
using uzkge::Limbs;
typedef uint64_t u64;

constexpr size_t N_WIRES = 5, N_SEL = 9, N_BLIND_SLOTS = 3, TAIL = 8;
const size_t HIDING[5] = {3, 3, 3, 2, 2};   // TurboCS::get_hiding_degree (turbo/mod.rs)

#define TRY(expr)                          \
    do {                                   \
        const int rc__ = (expr);           \
        if (rc__ != UZKGE_OK) return rc__; \
    } while (0)
#define CU(expr)                                                                                   \
    do {                                                                                           \
        const cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                                  \
            cudaGetLastError();                                                                    \
            return uz::api_fail(e__ == cudaErrorMemoryAllocation ? UZKGE_ERR_OOM : UZKGE_ERR_CUDA, \
                                (std::string("plonk prover: ") + cudaGetErrorString(e__)).c_str()); \
        }                                                                                          \
    } while (0)

// ---- scalars (Montgomery limbs on the host)
inline Limbs fr_sub(const Limbs& a, const Limbs& b) { return FR.add(a, FR.neg(b)); }
inline Limbs fr_pow_u64(Limbs a, u64 e) {
    Limbs acc = FR.one;
    while (e) {
        if (e & 1) acc = FR.mul(acc, a);
        a = FR.mul(a, a);
        e >>= 1;
    }
    return acc;
}
inline Limbs fr_from_u64(u64 v) { return FR.to_mont(Limbs{v, 0, 0, 0}); }
inline bool is_zero(const Limbs& a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
inline Limbs load(const u64* p) { return Limbs{p[0], p[1], p[2], p[3]}; }

// k Jacobian points (12 words) -> affine (8 words, zeros for the identity) with ONE field inversion (Montgomery's trick)
void batch_to_affine(const u64* jac, size_t k, u64* aff) {
    std::vector<Limbs> prefix(k);
    Limbs acc = FQ.one;
    for (size_t i = 0; i < k; i++) {
        const Limbs z = load(jac + 12 * i + 8);
        prefix[i] = acc;
        if (!is_zero(z)) acc = FQ.mul(acc, z);
    }
    Limbs inv = FQ.inverse(acc);
    for (size_t i = k; i-- > 0;) {
        const Limbs z = load(jac + 12 * i + 8);
        u64* out = aff + 8 * i;
        if (is_zero(z)) {
            memset(out, 0, 64);
            continue;
        }
        const Limbs zi = FQ.mul(inv, prefix[i]);     // 1 / z_i
        inv = FQ.mul(inv, z);
        const Limbs zi2 = FQ.mul(zi, zi);
        const Limbs x = FQ.mul(load(jac + 12 * i), zi2);
        const Limbs y = FQ.mul(load(jac + 12 * i + 4), FQ.mul(zi2, zi));
        memcpy(out, x.data(), 32);
        memcpy(out + 4, y.data(), 32);
    }
}

struct Poly {          // coefficients in HBM
    u64* p = nullptr;
    size_t len = 0;
};

struct Params {
    int device = 0;
    size_t n = 0, m = 0, factor = 0, num_vars = 0, n_public = 0, stride = 0;
    size_t ntt_batch = 1;      // vectors one transform call may carry (the scratch vector holds ntt_batch * m elements)
    bool shuffle = false;
    Limbs k[5], root, root_m, k1, k1_inv, anemoi_g, anemoi_ginv, edwards_a, z_h_inv[16];
    std::vector<void*> allocs;
    uint32_t *wiring = nullptr, *pub_rows = nullptr, *pub_wit = nullptr;
    u64 *group = nullptr, *coset_quotient = nullptr, *sigma = nullptr, *l1_coset = nullptr, *zero_poly = nullptr, *zero_coset = nullptr;
    Poly q[9], s[5], qb, prk[4], q_ecc, gen[12], pk[12];
    u64 *q_coset[9] = {}, *s_coset[5] = {}, *qb_coset = nullptr, *prk_coset[4] = {}, *q_ecc_coset = nullptr, *gen_coset[12] = {}, *pk_coset[12] = {};
    // per-proof workspace (allocated with the parameters; one proof at a time per handle)
    u64 *scratch = nullptr, *coset[8] = {}, *w_sel_coset[3] = {}, *wit = nullptr, *ext = nullptr, *sel_ev = nullptr, *z_ev = nullptr, *ztmp = nullptr;
    u64 *polys = nullptr;      // 16 buffers of stride elements: w[5], w_sel[3], z, t[5], r, (spare)
    u64 *pi = nullptr, *sh = nullptr, *q1 = nullptr, *q2 = nullptr, *lag_buf = nullptr, *small = nullptr;
    // device group (uzkge_cuda_plonk_params_upload_multi): one compact coset vector (n: the quotient map's output), the exchanged
    // quotient cosets (m elements, coset j at [j n, (j + 1) n))
    u64 *cbuf = nullptr, *tcos = nullptr;
    // a group member's own cosets j = rank mod G of the quotient domain in COMPACT form (n values each: element i = point j + factor i):
    // the preprocessed columns, gathered once from the size-m evaluations above, and this proof's wires / z / pi / witness selectors.
    // The quotient map of a coset then runs on contiguous vectors (the strided walk over the size-m arrays fetches 4x the bytes)
    struct CosetCols {
        size_t j = 0;
        u64 *q[9] = {}, *s[5] = {}, *qb = nullptr, *prk[4] = {}, *l1 = nullptr, *coset_quotient = nullptr, *q_ecc = nullptr, *gen[12] = {}, *pk[12] = {};
        u64* dyn = nullptr;      // 10 x n: w[5], w_sel[3], z, pi
    };
    std::vector<CosetCols> own_cosets;
    size_t group_rank = 0, group_size = 1;
    u64* pinned = nullptr;     // host, page-locked: results of the small device-to-host reads
    cudaStream_t st = nullptr, side = nullptr;
    cudaEvent_t ev = nullptr;
    std::mutex mu;

    u64* poly_buf(size_t i) const { return polys + i * stride * 4; }
};

std::mutex g_params_mu;
std::map<u64, std::unique_ptr<Params>> g_params;
u64 g_params_next = 1;
// a multi-device parameter handle (bit 62) owns one ordinary handle per member of the device group
constexpr u64 PARAMS_MULTI = 1ull << 62;
std::map<u64, std::vector<u64>> g_multi_params;

// ---- one proof on the whole device group: every member (one worker thread each) runs the prover below on replicated polynomials and
// meets the others only here.  The members' transcripts stay identical because every commitment is the sum of all partial sums.
using uz::SpinBarrier;
struct Group {
    size_t G = 1;
    SpinBarrier bar;
    std::vector<u64> partial;            // [2][G][16][12]: the members' partial sums of a commitment batch, double-buffered by batch parity
    std::vector<u64> evals;              // [32][4]: the round-4 evaluations, each computed by one member
    uz::GroupSrsParts srs, lag;          // the monomial SRS / the Lagrange commitment SRS, split over the members
    std::vector<Params*> params;         // the members' parameter sets (tcos: where a member receives the others' quotient cosets)
};

int dev_alloc(Params& P, size_t elems, u64** out, bool zero) {
    void* p = nullptr;
    CU(cudaMalloc(&p, (elems ? elems : 1) * 32));
    P.allocs.push_back(p);
    if (zero) CU(cudaMemsetAsync(p, 0, (elems ? elems : 1) * 32, P.st));
    *out = (u64*)p;
    return UZKGE_OK;
}

void release(Params& P) {
    cudaSetDevice(P.device);
    cudaDeviceSynchronize();
    for (void* p : P.allocs) cudaFree(p);
    if (P.pinned) cudaFreeHost(P.pinned);
    if (P.st) cudaStreamDestroy(P.st);
    if (P.side) cudaStreamDestroy(P.side);
    if (P.ev) cudaEventDestroy(P.ev);
}

// ---- device calls, all on the parameter set's stream
int ifft(Params& P, const u64* src, u64* out, size_t len_in) {
    return uzkge_cuda_ntt_fr_device(src, out, P.scratch, len_in, P.n, 1, nullptr, P.st);
}
int coset_fft(Params& P, const Poly& f, u64* out) {
    return uzkge_cuda_ntt_fr_device(f.p, out, P.scratch, f.len, P.m, 0, P.k1.data(), P.st);
}
// k transforms over one domain: batched launches where the workspace allows, else one by one
int ntt_many(Params& P, const u64* const* ins, u64* const* outs, const size_t* lens, size_t k, size_t size, int inverse, const u64* shift) {
    size_t batch = P.ntt_batch * P.m / size;       // the scratch vector holds ntt_batch * m elements
    if (batch > uz::NTT_MAX_BATCH) batch = uz::NTT_MAX_BATCH;
    if (P.ntt_batch == 1) batch = 1;               // large circuits: one vector fills the GPU, batching buys nothing
    for (size_t j0 = 0; j0 < k; j0 += batch) {
        const size_t kk = k - j0 < batch ? k - j0 : batch;
        if (kk == 1)
            TRY(uzkge_cuda_ntt_fr_device(ins[j0], outs[j0], P.scratch, lens[j0], size, inverse, shift, P.st));
        else
            TRY(uzkge_cuda_ntt_fr_batch_device((const void* const*)(ins + j0), (void* const*)(outs + j0), P.scratch, lens + j0, kk, size, inverse, shift, P.st));
    }
    return UZKGE_OK;
}
// coefficient form (host, `len` values) -> n zero-padded coefficients in HBM and the evaluations on the quotient coset
int preprocess(Params& P, const u64* host, size_t len, Poly* poly, u64** coset) {
    if (!host || len == 0) {   // the zero polynomial: every absent selector shares one buffer (and its evaluations hit L2)
        poly->p = P.zero_poly;
        poly->len = P.n;
        *coset = P.zero_coset;
        return UZKGE_OK;
    }
    if (len > P.n) return uz::api_fail(UZKGE_ERR_SIZE, "plonk_params_upload: a preprocessed polynomial has more than n coefficients");
    TRY(dev_alloc(P, P.n, &poly->p, len < P.n));
    poly->len = P.n;
    CU(cudaMemcpyAsync(poly->p, host, len * 32, cudaMemcpyHostToDevice, P.st));
    TRY(dev_alloc(P, P.m, coset, false));
    return coset_fft(P, *poly, *coset);
}

// compact copy of coset j of a size-m column (shared zero columns stay shared: one compact zero vector)
int coset_column(Params& P, const u64* col_m, size_t j, u64** out) {
    if (col_m == P.zero_coset) {
        *out = P.zero_coset;         // n zeros are a prefix of m zeros
        return UZKGE_OK;
    }
    if (!*out || *out == P.zero_coset) TRY(dev_alloc(P, P.n, out, false));
    return uzkge_cuda_fr_strided_copy_device(col_m, j, P.factor, *out, 0, 1, P.n, P.st);
}
// (re)build the compact public-key columns of a member's cosets (after uzkge_cuda_plonk_params_set_public_key)
int coset_public_key_columns(Params& P) {
    if (!P.shuffle) return UZKGE_OK;
    for (Params::CosetCols& c : P.own_cosets)
        for (int i = 0; i < 12; i++) TRY(coset_column(P, P.pk_coset[i], c.j, &c.pk[i]));
    return UZKGE_OK;
}
// every preprocessed column of member `rank`'s cosets, plus the per-proof buffers
int build_own_cosets(Params& P, size_t rank, size_t G) {
    P.group_rank = rank;
    P.group_size = G;
    for (size_t j = rank; j < P.factor; j += G) {
        Params::CosetCols c;
        c.j = j;
        for (int i = 0; i < 9; i++) TRY(coset_column(P, P.q_coset[i], j, &c.q[i]));
        for (int i = 0; i < 5; i++) TRY(coset_column(P, P.s_coset[i], j, &c.s[i]));
        TRY(coset_column(P, P.qb_coset, j, &c.qb));
        for (int i = 0; i < 4; i++) TRY(coset_column(P, P.prk_coset[i], j, &c.prk[i]));
        TRY(coset_column(P, P.l1_coset, j, &c.l1));
        TRY(coset_column(P, P.coset_quotient, j, &c.coset_quotient));
        if (P.shuffle) {
            TRY(coset_column(P, P.q_ecc_coset, j, &c.q_ecc));
            for (int i = 0; i < 12; i++) TRY(coset_column(P, P.gen_coset[i], j, &c.gen[i]));
        }
        TRY(dev_alloc(P, 10 * P.n, &c.dyn, true));       // pi stays zero without public inputs
        P.own_cosets.push_back(c);
    }
    TRY(coset_public_key_columns(P));
    CU(cudaStreamSynchronize(P.st));
    return UZKGE_OK;
}

struct Msm {           // one commitment to make: scalars in HBM
    const u64* p;
    size_t len;
};

}  // namespace

extern "C" {

UZKGE_API int32_t uzkge_cuda_srs_upload_lagrange_commit(const uint64_t* lagrange_xy, size_t n, const uint64_t* monomial_xy, size_t monomial_len,
                                                        uint32_t window_bits, uint64_t* handle) {
    if (!lagrange_xy || !monomial_xy || !handle) return uz::api_fail(UZKGE_ERR_ARG, "srs_upload_lagrange_commit: null pointer");
    if (n == 0 || monomial_len < n + N_BLIND_SLOTS) return uz::api_fail(UZKGE_ERR_SIZE, "srs_upload_lagrange_commit: the monomial SRS must hold n + 3 points");
    std::vector<u64> pts((n + 2 * N_BLIND_SLOTS) * 8);
    memcpy(pts.data(), lagrange_xy, n * 64);
    memcpy(pts.data() + n * 8, monomial_xy, N_BLIND_SLOTS * 64);
    memcpy(pts.data() + (n + N_BLIND_SLOTS) * 8, monomial_xy + n * 8, N_BLIND_SLOTS * 64);
    return uzkge_cuda_srs_upload(pts.data(), n + 2 * N_BLIND_SLOTS, window_bits, handle);
}

UZKGE_API int32_t uzkge_cuda_srs_upload_lagrange_commit_multi(const uint64_t* lagrange_xy, size_t n, const uint64_t* monomial_xy, size_t monomial_len,
                                                              uint32_t window_bits, uint64_t* handle) {
    if (!lagrange_xy || !monomial_xy || !handle) return uz::api_fail(UZKGE_ERR_ARG, "srs_upload_lagrange_commit_multi: null pointer");
    if (n == 0 || monomial_len < n + N_BLIND_SLOTS) return uz::api_fail(UZKGE_ERR_SIZE, "srs_upload_lagrange_commit_multi: the monomial SRS must hold n + 3 points");
    std::vector<u64> pts((n + 2 * N_BLIND_SLOTS) * 8);
    memcpy(pts.data(), lagrange_xy, n * 64);
    memcpy(pts.data() + n * 8, monomial_xy, N_BLIND_SLOTS * 64);
    memcpy(pts.data() + (n + N_BLIND_SLOTS) * 8, monomial_xy + n * 8, N_BLIND_SLOTS * 64);
    return uzkge_cuda_srs_upload_multi(pts.data(), n + 2 * N_BLIND_SLOTS, window_bits, UZKGE_MULTI_SPLIT, handle);
}

UZKGE_API int32_t uzkge_cuda_plonk_params_upload(const uzkge_plonk_params_desc* d, uint64_t* params_handle) {
    if (!d || !params_handle || !d->wiring || !d->permutation) return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_upload: null pointer");
    const size_t n = d->n, m = d->m;
    if (n < 2 || (n & (n - 1)) || m % n || m / n < 1 || m / n > 16 || n >= (1ull << 28))
        return uz::api_fail(UZKGE_ERR_SIZE, "plonk_params_upload: n must be a power of two, m a multiple of n with factor <= 16");
    if (d->n_public && (!d->public_vars_constraint_indices || !d->public_vars_witness_indices))
        return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_upload: null public-input indices");
    TRY(uzkge_cuda_init(-1));
    std::unique_ptr<Params> up(new Params());
    Params& P = *up;
    P.device = uzkge_cuda_get_device();
    CU(cudaSetDevice(P.device));
    P.n = n;
    P.m = m;
    P.factor = m / n;
    P.num_vars = d->num_vars;
    P.n_public = d->n_public;
    P.shuffle = d->shuffle != 0;
    P.stride = n + TAIL;
    for (int i = 0; i < 5; i++) P.k[i] = load(d->k[i]);
    P.anemoi_g = load(d->anemoi_generator);
    P.anemoi_ginv = load(d->anemoi_generator_inv);
    P.edwards_a = load(d->edwards_a);
    TRY(uzkge_cuda_fr_root_of_unity(n, P.root.data()));
    TRY(uzkge_cuda_fr_root_of_unity(m, P.root_m.data()));
    P.k1 = P.k[1];
    P.k1_inv = FR.inverse(P.k1);
    struct Guard {   // frees everything unless the upload completes
        Params* p;
        ~Guard() {
            if (p) release(*p);
        }
    } guard{&P};
    CU(cudaStreamCreateWithFlags(&P.st, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&P.side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&P.ev, cudaEventDisableTiming));
    CU(cudaHostAlloc((void**)&P.pinned, 4096, cudaHostAllocDefault));

    // workspace first (the transforms below need the scratch vector)
    // small circuits: the 5-8 transforms of a round travel in one launch per pass (uzkge_cuda_ntt_fr_batch_device)
    P.ntt_batch = n <= ((size_t)1 << 18) ? 8 : 1;
    TRY(dev_alloc(P, P.ntt_batch * m, &P.scratch, false));
    TRY(dev_alloc(P, n, &P.zero_poly, true));
    TRY(dev_alloc(P, m, &P.zero_coset, true));
    for (int i = 0; i < 8; i++) TRY(dev_alloc(P, m, &P.coset[i], i == 5));   // [5] = pi on the coset: stays zero without public inputs
    if (P.shuffle)
        for (int i = 0; i < 3; i++) TRY(dev_alloc(P, m, &P.w_sel_coset[i], false));
    TRY(dev_alloc(P, P.num_vars, &P.wit, false));
    TRY(dev_alloc(P, N_WIRES * P.stride, &P.ext, true));
    TRY(dev_alloc(P, 3 * P.stride, &P.sel_ev, true));
    TRY(dev_alloc(P, P.stride, &P.z_ev, true));
    TRY(dev_alloc(P, 4 * n, &P.ztmp, false));
    TRY(dev_alloc(P, 16 * P.stride, &P.polys, true));
    TRY(dev_alloc(P, n, &P.pi, true));
    TRY(dev_alloc(P, P.stride, &P.sh, false));
    TRY(dev_alloc(P, P.stride, &P.q1, false));
    TRY(dev_alloc(P, P.stride, &P.q2, false));
    TRY(dev_alloc(P, 5 * P.stride, &P.lag_buf, true));
    TRY(dev_alloc(P, 128, &P.small, true));

    // the domain and the quotient coset (indexer.rs:276-282)
    TRY(dev_alloc(P, n, &P.group, false));
    TRY(uzkge_cuda_fr_powers_device(P.root.data(), nullptr, n, P.group, P.st));
    TRY(dev_alloc(P, m, &P.coset_quotient, false));
    TRY(uzkge_cuda_fr_powers_device(P.root_m.data(), P.k1.data(), m, P.coset_quotient, P.st));
    // sigma = the permutation encoded into the group (indexer.rs:195-208): table[c n + i] = k_c w^i, sigma = table[perm]
    {
        u64* table;
        TRY(dev_alloc(P, N_WIRES * n, &table, false));
        for (size_t c = 0; c < N_WIRES; c++) TRY(uzkge_cuda_fr_powers_device(P.root.data(), P.k[c].data(), n, table + 4 * c * n, P.st));
        std::vector<uint32_t> perm32(N_WIRES * n);
        for (size_t i = 0; i < N_WIRES * n; i++) {
            if (d->permutation[i] >= N_WIRES * n) return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_upload: permutation entry out of range");
            perm32[i] = (uint32_t)d->permutation[i];
        }
        for (size_t i = 0; i < N_WIRES * n; i++)
            if (d->wiring[i] >= d->num_vars) return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_upload: wire index out of range");
        uint32_t* d_perm;
        CU(cudaMalloc((void**)&d_perm, N_WIRES * n * 4));
        P.allocs.push_back(d_perm);
        CU(cudaMemcpyAsync(d_perm, perm32.data(), N_WIRES * n * 4, cudaMemcpyHostToDevice, P.st));
        TRY(dev_alloc(P, N_WIRES * n, &P.sigma, false));
        TRY(uzkge_cuda_fr_gather_device(table, d_perm, N_WIRES * n, P.sigma, P.st));
        CU(cudaStreamSynchronize(P.st));   // perm32 goes out of scope
        CU(cudaMalloc((void**)&P.wiring, N_WIRES * n * 4));
        P.allocs.push_back(P.wiring);
        CU(cudaMemcpyAsync(P.wiring, d->wiring, N_WIRES * n * 4, cudaMemcpyHostToDevice, P.st));
    }
    if (P.n_public) {
        std::vector<uint32_t> rows(P.n_public), wi(P.n_public);
        for (size_t j = 0; j < P.n_public; j++) {
            if (d->public_vars_constraint_indices[j] >= n || d->public_vars_witness_indices[j] >= d->num_vars)
                return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_upload: public-input index out of range");
            rows[j] = (uint32_t)d->public_vars_constraint_indices[j];
            wi[j] = (uint32_t)d->public_vars_witness_indices[j];
        }
        CU(cudaMalloc((void**)&P.pub_rows, P.n_public * 4));
        P.allocs.push_back(P.pub_rows);
        CU(cudaMalloc((void**)&P.pub_wit, P.n_public * 4));
        P.allocs.push_back(P.pub_wit);
        CU(cudaMemcpyAsync(P.pub_rows, rows.data(), P.n_public * 4, cudaMemcpyHostToDevice, P.st));
        CU(cudaMemcpyAsync(P.pub_wit, wi.data(), P.n_public * 4, cudaMemcpyHostToDevice, P.st));
        CU(cudaStreamSynchronize(P.st));
    }
    // preprocessed polynomials and their coset evaluations
    for (size_t i = 0; i < N_SEL; i++) TRY(preprocess(P, d->q_polys[i], d->q_len[i], &P.q[i], &P.q_coset[i]));
    for (size_t i = 0; i < N_WIRES; i++) TRY(preprocess(P, d->s_polys[i], d->s_len[i], &P.s[i], &P.s_coset[i]));
    TRY(preprocess(P, d->qb_poly, d->qb_len, &P.qb, &P.qb_coset));
    for (size_t i = 0; i < 4; i++) TRY(preprocess(P, d->q_prk_polys[i], d->q_prk_len[i], &P.prk[i], &P.prk_coset[i]));
    if (P.shuffle) {
        TRY(preprocess(P, d->q_ecc_poly, d->q_ecc_len, &P.q_ecc, &P.q_ecc_coset));
        for (size_t i = 0; i < 12; i++) TRY(preprocess(P, d->q_shuffle_generator_polys[i], d->gen_len[i], &P.gen[i], &P.gen_coset[i]));
        for (size_t i = 0; i < 12; i++) TRY(preprocess(P, d->q_shuffle_public_key_polys[i], d->pk_len[i], &P.pk[i], &P.pk_coset[i]));
    }
    // L1 on the coset: l1_coefs = (1 + X + ... + X^(n-1)) / n  <=>  the evaluation vector (1, 0, ..., 0) (indexer.rs:343-346)
    {
        Poly l1;
        TRY(dev_alloc(P, n, &l1.p, true));
        l1.len = n;
        const size_t idx0 = 0;
        const Limbs nn = fr_from_u64((u64)n);
        TRY(uzkge_cuda_fr_add_sparse_device(l1.p, &idx0, nn.data(), 1, P.st));      // plonk.py: l1 evals = [n, 0, ...]
        TRY(ifft(P, l1.p, l1.p, n));
        TRY(dev_alloc(P, m, &P.l1_coset, false));
        TRY(coset_fft(P, l1, P.l1_coset));
    }
    // Z_H^-1 on the coset takes `factor` values (helpers.rs:244-253)
    {
        Limbs mult = fr_pow_u64(P.k1, (u64)n);
        const Limbs step = fr_pow_u64(P.root_m, (u64)n);
        for (size_t j = 0; j < P.factor; j++) {
            P.z_h_inv[j] = FR.inverse(fr_sub(mult, FR.one));
            mult = FR.mul(mult, step);
        }
    }
    CU(cudaStreamSynchronize(P.st));
    guard.p = nullptr;
    std::lock_guard<std::mutex> lock(g_params_mu);
    const u64 h = g_params_next++;
    g_params[h] = std::move(up);
    *params_handle = h;
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_plonk_params_set_public_key(uint64_t params_handle, const uint64_t* const polys[12], const size_t len[12]) {
    if (!polys || !len) return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_set_public_key: null pointer");
    if (params_handle & PARAMS_MULTI) {
        std::vector<u64> subs;
        {
            std::lock_guard<std::mutex> lock(g_params_mu);
            auto it = g_multi_params.find(params_handle);
            if (it == g_multi_params.end()) return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_params_set_public_key: unknown handle");
            subs = it->second;
        }
        for (u64 h : subs) TRY(uzkge_cuda_plonk_params_set_public_key(h, polys, len));
        return UZKGE_OK;
    }
    Params* Pp;
    {
        std::lock_guard<std::mutex> lock(g_params_mu);
        auto it = g_params.find(params_handle);
        if (it == g_params.end()) return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_params_set_public_key: unknown handle");
        Pp = it->second.get();
    }
    Params& P = *Pp;
    std::lock_guard<std::mutex> lock(P.mu);
    if (!P.shuffle) return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_set_public_key: the parameters were built without the shuffle feature set");
    TRY(uzkge_cuda_set_device(P.device));
    for (size_t i = 0; i < 12; i++) {
        if (len[i] > P.n) return uz::api_fail(UZKGE_ERR_SIZE, "plonk_params_set_public_key: more than n coefficients");
        if (P.pk[i].p == P.zero_poly) {        // was a shared zero: give it its own storage
            TRY(dev_alloc(P, P.n, &P.pk[i].p, false));
            TRY(dev_alloc(P, P.m, &P.pk_coset[i], false));
        }
        CU(cudaMemsetAsync(P.pk[i].p, 0, P.n * 32, P.st));
        if (len[i]) CU(cudaMemcpyAsync(P.pk[i].p, polys[i], len[i] * 32, cudaMemcpyHostToDevice, P.st));
        TRY(coset_fft(P, P.pk[i], P.pk_coset[i]));
    }
    TRY(coset_public_key_columns(P));      // a group member keeps compact copies of its cosets' columns
    CU(cudaStreamSynchronize(P.st));
    return UZKGE_OK;
}

UZKGE_API int32_t uzkge_cuda_plonk_params_free(uint64_t params_handle) {
    if (params_handle & PARAMS_MULTI) {
        std::vector<u64> subs;
        {
            std::lock_guard<std::mutex> lock(g_params_mu);
            auto it = g_multi_params.find(params_handle);
            if (it == g_multi_params.end()) return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_params_free: unknown handle");
            subs = it->second;
            g_multi_params.erase(it);
        }
        int rc = UZKGE_OK;
        for (u64 h : subs) {
            const int r = uzkge_cuda_plonk_params_free(h);
            if (r != UZKGE_OK) rc = r;
        }
        return rc;
    }
    std::unique_ptr<Params> victim;
    {
        std::lock_guard<std::mutex> lock(g_params_mu);
        auto it = g_params.find(params_handle);
        if (it == g_params.end()) return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_params_free: unknown handle");
        victim = std::move(it->second);
        g_params.erase(it);
    }
    std::lock_guard<std::mutex> lock(victim->mu);
    release(*victim);
    return UZKGE_OK;
}

}  // extern "C"

namespace {

// One proof.  grp == nullptr: the whole proof on the parameter set's device.  Otherwise this is member `rank` of a device group: the
// same code on replicated data, except that a commitment covers only the member's slice of the SRS (the partial sums meet in
// grp->partial) and the quotient is evaluated only on the member's cosets of the quotient domain (exchanged through Params::tcos).
int prove_impl(const uzkge_plonk_prove_args* a, uzkge_plonk_proof* proof, Group* grp, size_t rank) {
    Params* Pp;
    {
        std::lock_guard<std::mutex> lock(g_params_mu);
        auto it = g_params.find(a->params);
        if (it == g_params.end()) return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_prove: unknown parameter handle");
        Pp = it->second.get();
    }
    Params& P = *Pp;
    std::lock_guard<std::mutex> lock(P.mu);
    TRY(uzkge_cuda_set_device(P.device));
    const size_t G = grp ? grp->G : 1;
    size_t commit_seq = 0;
    const size_t n = P.n, m = P.m, stride = P.stride;
    const bool shuffle = P.shuffle;
    const bool lagrange = a->lagrange_srs != 0;
    const bool lagrange_all = lagrange && a->lagrange_all != 0;
    const size_t need_blinds = 13 + (shuffle ? 6 : 0) + 3 + 5;
    if (a->n_blinds < need_blinds) return uz::api_fail(UZKGE_ERR_SIZE, "plonk_prove: 21 blinds are needed (27 with the shuffle feature set)");
    // SRS sizes (the reference's DegreeError, kzg_poly_commitment.rs:283-285, is checked before every commitment)
    size_t srs_n = 0, lag_n = 0;
    uzkge_srs_info info;
    if (a->srs) {
        if (grp) {
            srs_n = grp->srs.n;       // resolved by the caller (the group call holds the registry's lock)
        } else {
            TRY(uzkge_cuda_srs_info(a->srs, &info));
            srs_n = info.n;
        }
    }
    if (lagrange) {
        if (grp) {
            lag_n = grp->lag.n;
        } else {
            TRY(uzkge_cuda_srs_info(a->lagrange_srs, &info));
            lag_n = info.n;
        }
        if (lag_n != n + 2 * N_BLIND_SLOTS) return uz::api_fail(UZKGE_ERR_SIZE, "plonk_prove: the Lagrange commitment SRS does not match the circuit size");
    }
    if (!lagrange_all && srs_n < n + 3) return uz::api_fail(UZKGE_ERR_SIZE, "plonk_prove: the monomial SRS must hold n + 3 points");
    memset(proof, 0, sizeof *proof);
    const uint64_t launches0 = uzkge_cuda_launch_count();
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](int stage) {
        const auto now = std::chrono::steady_clock::now();
        proof->rounds_ms[stage] += std::chrono::duration<double, std::milli>(now - t_prev).count();
        t_prev = now;
    };
    uzkge::Transcript tr = uzkge::Transcript::from_state(a->transcript, a->transcript_len);
    const u64* blind = a->blinds;
    auto next_blind = [&]() {
        const Limbs b = load(blind);
        blind += 4;
        return b;
    };
    cudaStream_t st = P.st;

    // ---- helpers over this proof's buffers
    // sparse updates of a round (blinds) are collected and applied in one launch
    struct Sparse {
        std::vector<void*> p;
        std::vector<size_t> idx;
        std::vector<u64> vals;
        void add(u64* poly, size_t i, const Limbs& v) {
            p.push_back(poly);
            idx.push_back(i);
            vals.insert(vals.end(), v.begin(), v.end());
        }
    } sparse;
    auto flush_sparse = [&]() -> int {
        for (size_t j0 = 0; j0 < sparse.p.size(); j0 += UZKGE_SPARSE_MULTI_MAX) {
            const size_t kk = sparse.p.size() - j0 < UZKGE_SPARSE_MULTI_MAX ? sparse.p.size() - j0 : UZKGE_SPARSE_MULTI_MAX;
            TRY(uzkge_cuda_fr_add_sparse_multi_device(sparse.p.data() + j0, sparse.idx.data() + j0, sparse.vals.data() + 4 * j0, kk, st));
        }
        sparse = Sparse();
        return UZKGE_OK;
    };
    // hide_polynomial (helpers.rs:139-154): f += (b_0 + b_1 X + ...) (X^zeroing_degree - 1); returns the blinds (queued: flush_sparse)
    // group members on large circuits DEAL the interpolations of a round (job j to member j mod G) and pull the others' results
    // over peer memory; a dealt polynomial is blinded by its owner only.  owner_of_buf[b]: the member that produces poly_buf(b), -1 = all
    const bool deal = grp && G > 1 && n >= ((size_t)1 << uz::g_group_deal_min_log_n);
    int owner_of_buf[16];
    for (int b = 0; b < 16; b++) owner_of_buf[b] = -1;
    auto hide = [&](Poly& f, size_t hiding, std::vector<Limbs>* out) -> int {
        const size_t buf = (size_t)(f.p - P.polys) / (4 * stride);
        const bool apply = !deal || buf >= 16 || owner_of_buf[buf] < 0 || (size_t)owner_of_buf[buf] == rank;
        for (size_t i = 0; i < hiding; i++) {
            const Limbs b = next_blind();
            out->push_back(b);
            if (!apply) continue;
            sparse.add(f.p, i, b);
            sparse.add(f.p, n + i, FR.neg(b));
        }
        if (f.len < n + hiding) f.len = n + hiding;
        return UZKGE_OK;
    };
    // slots [first, first + 6) <- [b_0 b_1 b_2 | -b_0 -b_1 -b_2] (missing blinds are zero): the blind terms of a Lagrange commitment.
    // The caller zeroes the slots first; the values travel as kernel arguments (a copy from pageable memory would synchronise)
    auto set_blind_slots = [&](u64* buf, size_t first, const std::vector<Limbs>& blinds) -> int {
        for (size_t i = 0; i < blinds.size(); i++) {
            sparse.add(buf, first + i, blinds[i]);
            sparse.add(buf, first + N_BLIND_SLOTS + i, FR.neg(blinds[i]));
        }
        return UZKGE_OK;
    };
    // zero the 8-element tails that follow the n values of `count` consecutive stride-spaced vectors
    auto zero_tails = [&](u64* base, size_t count) -> int {
        CU(cudaMemset2DAsync(base + 4 * n, stride * 32, 0, TAIL * 32, count, st));
        return UZKGE_OK;
    };
    // commit k vectors over one SRS; `overlap` (device work that does not depend on them) is enqueued behind the MSMs, the results are
    // read back on the side stream so that the host can hash while the GPU keeps working.  out_aff: k x 8 words (affine, Montgomery)
    auto commit = [&](u64 handle, const size_t handle_n, const std::vector<Msm>& v, u64* out_aff, const std::function<int()>& overlap) -> int {
        const size_t k = v.size();
        for (const Msm& x : v)
            if (x.len > handle_n) return uz::api_fail(UZKGE_ERR_SIZE, "plonk_prove: DegreeError (a polynomial does not fit the SRS: unsatisfied witness?)");
        u64* d_out = P.small;                     // 16 x 12 words
        const void* ptrs[16];
        size_t lens[16];
        if (k > 16) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: more than 16 commitments in one batch");
        // a group member commits the part of every vector that falls into its slice [lo, hi) of the SRS
        size_t lo = 0, hi = handle_n;
        bool mine = true;
        if (grp) {
            const uz::GroupSrsParts& parts = (lagrange && handle == a->lagrange_srs) ? grp->lag : grp->srs;
            mine = rank < parts.sub.size();
            if (mine) {
                handle = parts.sub[rank];
                lo = parts.lo[rank];
                hi = parts.hi[rank];
            }
        }
        for (size_t j = 0; j < k; j++) {
            const size_t b = v[j].len < hi ? v[j].len : hi;
            lens[j] = b > lo ? b - lo : 0;
            ptrs[j] = v[j].p + 4 * (lens[j] ? lo : 0);
        }
        if (!mine)
            CU(cudaMemsetAsync(d_out, 0, k * 96, st));      // Z = 0: the identity
        else if (k == 1)
            TRY(uzkge_cuda_msm_g1_device(handle, 0, ptrs[0], lens[0], d_out, st));
        else
            TRY(uzkge_cuda_msm_g1_batch_device(handle, 0, ptrs, lens, k, d_out, st));
        proof->msm += (uint32_t)k;
        CU(cudaEventRecord(P.ev, st));
        if (overlap) TRY(overlap());
        CU(cudaStreamWaitEvent(P.side, P.ev, 0));
        CU(cudaMemcpyAsync(P.pinned, d_out, k * 96, cudaMemcpyDeviceToHost, P.side));
        CU(cudaStreamSynchronize(P.side));
        // the next user of P.small is ordered after this read: it is enqueued on `st` only after the host has the results
        if (!grp) {
            batch_to_affine(P.pinned, k, out_aff);
            return UZKGE_OK;
        }
        // "the partial G1 sums are combined with one projective add each": every member adds all G partial sums, in member order
        u64* slots = grp->partial.data() + (commit_seq++ & 1) * G * 16 * 12;
        memcpy(slots + rank * 16 * 12, P.pinned, k * 96);
        if (!grp->bar.wait()) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: another member of the device group failed");
        u64 sums[16 * 12], col[uz::UZ_MAX_DEVICES * 12];
        for (size_t j = 0; j < k; j++) {
            for (size_t r = 0; r < G; r++) memcpy(col + 12 * r, slots + (r * 16 + j) * 12, 96);
            uz::group_sum_jacobians(col, G, sums + 12 * j);
        }
        batch_to_affine(sums, k, out_aff);
        return UZKGE_OK;
    };
    // coefficient vectors of up to n + 3 entries over the Lagrange bases (helpers.rs:1363-1391, pcs.rs:139-163): with f = f_lo + X^n f_hi,
    // f(tau) G = MSM(L_i(tau) G, f_lo on H) + sum_i f_hi[i] SRS[n + i] -- one forward transform and one MSM over [f_lo on H | 0 0 0 | f_hi]
    auto commit_coefs_lagrange = [&](const std::vector<Poly>& polys, u64* out_aff) -> int {
        std::vector<Msm> v;
        const u64* ins[8];
        u64* outs[8];
        size_t lens[8], k = 0;
        for (size_t i = 0; i < polys.size(); i++) {
            const Poly& f = polys[i];
            if (f.len > n + N_BLIND_SLOTS) return uz::api_fail(UZKGE_ERR_SIZE, "plonk_prove: DegreeError (a polynomial does not fit the SRS: unsatisfied witness?)");
            u64* buf = P.lag_buf + 4 * i * stride;
            CU(cudaMemsetAsync(buf + 4 * n, 0, TAIL * 32, st));
            if (f.len == 0) {
                CU(cudaMemsetAsync(buf, 0, n * 32, st));
            } else {
                ins[k] = f.p;
                outs[k] = buf;
                lens[k++] = f.len < n ? f.len : n;
                if (f.len > n)
                    CU(cudaMemcpyAsync(buf + 4 * (n + N_BLIND_SLOTS), f.p + 4 * n, (f.len - n) * 32, cudaMemcpyDeviceToDevice, st));
            }
            v.push_back({buf, n + 2 * N_BLIND_SLOTS});
        }
        TRY(ntt_many(P, ins, outs, lens, k, n, 0, nullptr));
        proof->fft_n += (uint32_t)k;
        return commit(a->lagrange_srs, lag_n, v, out_aff, nullptr);
    };
    auto append_commitments = [&](const u64* aff, size_t k) {
        for (size_t j = 0; j < k; j++) {
            std::array<u64, 8> pt;
            memcpy(pt.data(), aff + 8 * j, 64);
            tr.append_commitment(pt);
        }
    };

    // ---- group members: the quotient round by cosets.  The m = factor * n points k[1] w_m^p are the cosets g_j <w_n>, g_j = k[1] w_m^j,
    // p = factor * i + j.  On coset j a polynomial of n + 3 coefficients is a size-n coset transform of its coefficients folded with
    // X^n = g_j^n, and the map of a coset needs nothing from the other cosets: member r takes the cosets j = r mod G
    if (grp && (P.group_rank != rank || P.group_size != G))
        return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_prove: the parameter set was not uploaded for this member of the device group");
    // fs[i] on this member's cosets, into slot slots[i] of the coset's compact buffers (Params::CosetCols::dyn: w0..w4, w_sel0..2, z, pi)
    enum { SLOT_W = 0, SLOT_W_SEL = 5, SLOT_Z = 8, SLOT_PI = 9 };
    auto eval_on_my_cosets = [&](const Poly* fs, const size_t* slots, size_t k) -> int {
        if (k > 10) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: more than 10 polynomials per coset pass");
        for (const Params::CosetCols& c : P.own_cosets) {
            const Limbs g = FR.mul(P.k1, fr_pow_u64(P.root_m, (u64)c.j));
            const Limbs gn = fr_pow_u64(g, (u64)n);
            u64 fold[8];
            memcpy(fold, FR.one.data(), 32);
            memcpy(fold + 4, gn.data(), 32);
            const u64* ins[10];
            u64* cs[10];
            size_t lens[10];
            for (size_t i = 0; i < k; i++) {
                // the transform runs in place on a COPY of the first n coefficients with the tail folded in (head += g^n * tail): the
                // polynomial itself is never touched -- other members of the group may be reading it over peer memory
                const size_t extra = fs[i].len > n ? fs[i].len - n : 0;
                if (extra > TAIL) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: a polynomial exceeds n + 8 coefficients");
                cs[i] = c.dyn + 4 * n * slots[i];
                lens[i] = fs[i].len < n ? (fs[i].len ? fs[i].len : 1) : n;
                CU(cudaMemcpyAsync(cs[i], fs[i].p, lens[i] * 32, cudaMemcpyDeviceToDevice, st));
                if (extra) {
                    const void* two[2] = {cs[i], fs[i].p + 4 * n};
                    const size_t two_len[2] = {extra, extra};
                    TRY(uzkge_cuda_fr_lincomb_device(two, two_len, fold, 2, cs[i], extra, st));
                }
                ins[i] = cs[i];
            }
            TRY(ntt_many(P, ins, cs, lens, k, n, 0, g.data()));
            proof->fft_n += (uint32_t)k;
        }
        return UZKGE_OK;
    };

    // ---- 0. witness into HBM; 1. the PI polynomial (helpers.rs:111-131)
    if (a->witness_on_device) {
        CU(cudaMemcpyAsync(P.wit, a->witness, P.num_vars * 32, grp ? cudaMemcpyDefault : cudaMemcpyDeviceToDevice, st));
    } else if (grp && G > 1) {
        // a host witness crosses the host links once in total: member r uploads slice r, then pulls the other slices from its peers
        auto lo_of = [&](size_t r) { return r * P.num_vars / G; };
        const size_t lo = lo_of(rank), hi = lo_of(rank + 1);
        if (hi > lo) CU(cudaMemcpyAsync(P.wit + 4 * lo, a->witness + 4 * lo, (hi - lo) * 32, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        if (!grp->bar.wait()) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: another member of the device group failed");
        for (size_t r = 0; r < G; r++) {
            if (r == rank) continue;
            const size_t a0 = lo_of(r), a1 = lo_of(r + 1);
            Params* Q = grp->params[r];
            if (a1 > a0) CU(cudaMemcpyPeerAsync(P.wit + 4 * a0, P.device, Q->wit + 4 * a0, Q->device, (a1 - a0) * 32, st));
        }
    } else {
        CU(cudaMemcpyAsync(P.wit, a->witness, P.num_vars * 32, cudaMemcpyHostToDevice, st));
    }
    Poly pi{P.pi, n};
    if (P.n_public) {
        CU(cudaMemsetAsync(P.pi, 0, n * 32, st));
        TRY(uzkge_cuda_fr_gather_scatter_device(P.wit, P.pub_wit, P.pi, P.pub_rows, P.n_public, st));
        TRY(ifft(P, P.pi, P.pi, n));
        proof->ifft_n++;
    }
    // every polynomial buffer of this proof starts as zeros beyond what is written below (16 x stride elements: the tails matter)
    CU(cudaMemsetAsync(P.polys, 0, 16 * stride * 32, st));

    // ---- 2. wire polynomials: extend the witness, interpolate, hide (prover.rs:148-176)
    Poly w_polys[5], w_sel_polys[3], z_poly, t_polys[5], r_poly;
    std::vector<Limbs> w_blinds[5], w_sel_blinds[3], z_blinds;
    const bool has_remark = shuffle && a->w_sel_evals[0] && a->w_sel_evals[1] && a->w_sel_evals[2];
    {
        const u64* ins[8];
        u64* outs[8];
        size_t lens[8], k = 0;
        for (size_t i = 0; i < N_WIRES; i++) {
            u64* ev = P.ext + 4 * i * stride;
            TRY(uzkge_cuda_fr_gather_device(P.wit, P.wiring + i * n, n, ev, st));
            w_polys[i] = {P.poly_buf(i), n};
            ins[k] = ev;
            outs[k] = w_polys[i].p;
            lens[k++] = n;
        }
        // ---- 3. (`shuffle`) witness-selector polynomials (prover.rs:177-191): interpolated with the wires, hidden after them
        if (shuffle)
            for (size_t i = 0; i < 3; i++) {
                u64* ev = P.sel_ev + 4 * i * stride;
                w_sel_polys[i] = {P.poly_buf(5 + i), 1};
                if (has_remark) {
                    CU(cudaMemcpyAsync(ev, a->w_sel_evals[i], n * 32, cudaMemcpyHostToDevice, st));
                    w_sel_polys[i].len = n;
                    ins[k] = ev;
                    outs[k] = w_sel_polys[i].p;
                    lens[k++] = n;
                } else {
                    CU(cudaMemsetAsync(ev, 0, n * 32, st));
                }
            }
        if (deal) {
            const u64* oins[8];
            u64* oouts[8];
            size_t olens[8], ok = 0;
            for (size_t j = 0; j < k; j++) {
                owner_of_buf[(size_t)(outs[j] - P.polys) / (4 * stride)] = (int)(j % G);
                if (j % G != rank) continue;
                oins[ok] = ins[j];
                oouts[ok] = outs[j];
                olens[ok++] = lens[j];
            }
            TRY(ntt_many(P, oins, oouts, olens, ok, n, 1, nullptr));
        } else {
            TRY(ntt_many(P, ins, outs, lens, k, n, 1, nullptr));
        }
        proof->ifft_n += (uint32_t)k;
        for (size_t i = 0; i < N_WIRES; i++) TRY(hide(w_polys[i], HIDING[i], &w_blinds[i]));     // the RNG order: wires, then selectors
        if (shuffle)
            for (size_t i = 0; i < 3; i++) TRY(hide(w_sel_polys[i], 2, &w_sel_blinds[i]));
        TRY(flush_sparse());
        if (deal) {
            // every member's polynomials (blinds included) are complete: pull the others' over NVLink.  Nobody writes a wire
            // polynomial after this point, so no second barrier is needed
            CU(cudaStreamSynchronize(st));
            if (!grp->bar.wait()) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: another member of the device group failed");
            for (size_t j = 0; j < k; j++) {
                if (j % G == rank) continue;
                Params* Q = grp->params[j % G];
                const size_t buf = (size_t)(outs[j] - P.polys) / (4 * stride);
                CU(cudaMemcpyPeerAsync(P.poly_buf(buf), P.device, Q->poly_buf(buf), Q->device, stride * 32, st));
            }
        }
    }
    // the quotient round's coset evaluations of these polynomials depend on no challenge: they run behind the MSMs
    auto wire_cosets = [&]() -> int {
        if (grp) {
            Poly fs[8];
            size_t slots[8], kk = 0;
            for (size_t i = 0; i < N_WIRES; i++) {
                fs[kk] = w_polys[i];
                slots[kk++] = SLOT_W + i;
            }
            if (shuffle)
                for (size_t i = 0; i < 3; i++) {
                    fs[kk] = w_sel_polys[i];
                    slots[kk++] = SLOT_W_SEL + i;
                }
            return eval_on_my_cosets(fs, slots, kk);
        }
        const u64* ins[8];
        u64* outs[8];
        size_t lens[8], k = 0;
        for (size_t i = 0; i < N_WIRES; i++) {
            ins[k] = w_polys[i].p;
            outs[k] = P.coset[i];
            lens[k++] = w_polys[i].len;
        }
        if (shuffle)
            for (size_t i = 0; i < 3; i++) {
                ins[k] = w_sel_polys[i].p;
                outs[k] = P.w_sel_coset[i];
                lens[k++] = w_sel_polys[i].len;
            }
        proof->coset_fft_m += (uint32_t)k;
        return ntt_many(P, ins, outs, lens, k, m, 0, P.k1.data());
    };
    {
        std::vector<Msm> wires, sels;
        u64 wire_scheme = a->srs, sel_scheme = a->srs;
        size_t wire_n = srs_n, sel_n = srs_n;
        if (lagrange) {
            TRY(zero_tails(P.ext, N_WIRES));
            for (size_t i = 0; i < N_WIRES; i++) {
                TRY(set_blind_slots(P.ext, i * stride + n, w_blinds[i]));
                wires.push_back({P.ext + 4 * i * stride, n + 2 * N_BLIND_SLOTS});
            }
            wire_scheme = a->lagrange_srs;
            wire_n = lag_n;
        } else {
            for (size_t i = 0; i < N_WIRES; i++) wires.push_back({w_polys[i].p, w_polys[i].len});
        }
        if (shuffle) {
            if (lagrange_all) {
                TRY(zero_tails(P.sel_ev, 3));
                for (size_t i = 0; i < 3; i++) {
                    TRY(set_blind_slots(P.sel_ev, i * stride + n, w_sel_blinds[i]));
                    sels.push_back({P.sel_ev + 4 * i * stride, n + 2 * N_BLIND_SLOTS});
                }
                sel_scheme = a->lagrange_srs;
                sel_n = lag_n;
            } else {
                for (size_t i = 0; i < 3; i++) sels.push_back({w_sel_polys[i].p, w_sel_polys[i].len});
            }
        }
        TRY(flush_sparse());
        u64 aff[8 * 8];
        if (shuffle && sel_scheme == wire_scheme) {
            std::vector<Msm> all = wires;
            all.insert(all.end(), sels.begin(), sels.end());
            TRY(commit(wire_scheme, wire_n, all, aff, wire_cosets));
        } else {
            TRY(commit(wire_scheme, wire_n, wires, aff, wire_cosets));
            if (shuffle) TRY(commit(sel_scheme, sel_n, sels, aff + 8 * N_WIRES, nullptr));
        }
        memcpy(proof->cm_w, aff, sizeof proof->cm_w);
        if (shuffle) memcpy(proof->cm_w_sel, aff + 8 * N_WIRES, sizeof proof->cm_w_sel);
        append_commitments(aff, N_WIRES + (shuffle ? 3 : 0));
    }
    mark(0);

    // ---- 4. beta, gamma;  5. z: running product on H, interpolate, hide, commit (helpers.rs:160-220)
    const Limbs beta = tr.get_challenge_field_elem();
    tr.append_single_byte(0x01);
    const Limbs gamma = tr.get_challenge_field_elem();
    {
        const void* d_w[5];
        const void* d_sigma[5];
        for (size_t i = 0; i < N_WIRES; i++) {
            d_w[i] = P.ext + 4 * i * stride;
            d_sigma[i] = P.sigma + 4 * i * n;
        }
        u64 kk[20];
        for (int i = 0; i < 5; i++) memcpy(kk + 4 * i, P.k[i].data(), 32);
        TRY(uzkge_cuda_plonk_z_evals_fr_device(d_w, d_sigma, P.group, kk, beta.data(), gamma.data(), n, P.z_ev, P.ztmp, st));
        z_poly = {P.poly_buf(8), n};
        TRY(ifft(P, P.z_ev, z_poly.p, n));
        proof->ifft_n++;
        TRY(hide(z_poly, 3, &z_blinds));
        if (lagrange) {
            TRY(zero_tails(P.z_ev, 1));
            TRY(set_blind_slots(P.z_ev, n, z_blinds));
        }
        TRY(flush_sparse());
        auto z_coset = [&]() -> int {
            if (grp) {
                const size_t slot = SLOT_Z;
                return eval_on_my_cosets(&z_poly, &slot, 1);
            }
            proof->coset_fft_m++;
            return coset_fft(P, z_poly, P.coset[6]);
        };
        u64 aff[8];
        if (lagrange) {
            TRY(commit(a->lagrange_srs, lag_n, {{P.z_ev, n + 2 * N_BLIND_SLOTS}}, aff, z_coset));
        } else {
            TRY(commit(a->srs, srs_n, {{z_poly.p, z_poly.len}}, aff, z_coset));
        }
        memcpy(proof->cm_z, aff, 64);
        append_commitments(aff, 1);
    }
    mark(1);

    // ---- 6. alpha;  7. t = numerator / Z_H on the coset k[1] <w_m>, back to coefficients (helpers.rs:223-678)
    const Limbs alpha = tr.get_challenge_field_elem();
    if (P.n_public) {
        if (grp) {
            const size_t slot = SLOT_PI;
            TRY(eval_on_my_cosets(&pi, &slot, 1));
        } else {
            TRY(coset_fft(P, pi, P.coset[5]));
            proof->coset_fft_m++;
        }
    }
    u64* t_buf = P.coset[7];
    {
        uzkge_quotient_args qa;
        memset(&qa, 0, sizeof qa);
        for (size_t i = 0; i < N_WIRES; i++) qa.w[i] = P.coset[i];
        for (size_t i = 0; i < N_SEL; i++) qa.q[i] = P.q_coset[i];
        qa.pi = P.coset[5];
        qa.z = P.coset[6];
        for (size_t i = 0; i < N_WIRES; i++) qa.s[i] = P.s_coset[i];
        qa.coset_quotient = P.coset_quotient;
        qa.l1 = P.l1_coset;
        qa.qb = P.qb_coset;
        for (size_t i = 0; i < 4; i++) qa.q_prk[i] = P.prk_coset[i];
        for (int i = 0; i < 5; i++) memcpy(qa.k[i], P.k[i].data(), 32);
        memcpy(qa.alpha, alpha.data(), 32);
        memcpy(qa.beta, beta.data(), 32);
        memcpy(qa.gamma, gamma.data(), 32);
        memcpy(qa.anemoi_generator, P.anemoi_g.data(), 32);
        memcpy(qa.anemoi_generator_inv, P.anemoi_ginv.data(), 32);
        for (size_t j = 0; j < P.factor; j++) memcpy(qa.z_h_inv[j], P.z_h_inv[j].data(), 32);
        qa.m = m;
        qa.factor = P.factor;
        if (shuffle) {
            uzkge_quotient_shuffle_args sa;
            memset(&sa, 0, sizeof sa);
            for (size_t i = 0; i < 3; i++) sa.w_sel[i] = P.w_sel_coset[i];
            sa.q_ecc = P.q_ecc_coset;
            for (size_t i = 0; i < 12; i++) {
                sa.pk[i] = P.pk_coset[i];
                sa.gen[i] = P.gen_coset[i];
            }
            memcpy(sa.edwards_a, P.edwards_a.data(), 32);
            if (!grp) TRY(uzkge_cuda_plonk_quotient_shuffle_fr_device(&qa, &sa, t_buf, st));
        } else if (!grp) {
            TRY(uzkge_cuda_plonk_quotient_fr_device(&qa, t_buf, st));
        }
        if (grp) {
            // the map on each of this member's cosets, over the coset's COMPACT columns: a domain of n points with factor 1 (the
            // omega-shifted point is the next element, Z_H^-1 is the coset's constant); then back to coefficients, coset by coset:
            // the size-n coset iFFT of the values on coset j is u_j[r] = T_r(g_j^n), where t(X) = sum_r X^r T_r(X^n).  Every member
            // sends its u_j (n values) into every other member's Params::tcos over peer memory; a factor-point inverse DFT per r
            // over the cosets then gives t's coefficients -- no transform of the whole 6n domain, nothing but the cosets on NVLink
            for (const Params::CosetCols& c : P.own_cosets) {
                uzkge_quotient_args qc = qa;
                for (size_t i = 0; i < N_WIRES; i++) {
                    qc.w[i] = c.dyn + 4 * n * (SLOT_W + i);
                    qc.s[i] = c.s[i];
                }
                for (size_t i = 0; i < N_SEL; i++) qc.q[i] = c.q[i];
                qc.pi = c.dyn + 4 * n * SLOT_PI;
                qc.z = c.dyn + 4 * n * SLOT_Z;
                qc.coset_quotient = c.coset_quotient;
                qc.l1 = c.l1;
                qc.qb = c.qb;
                for (size_t i = 0; i < 4; i++) qc.q_prk[i] = c.prk[i];
                memset(qc.z_h_inv, 0, sizeof qc.z_h_inv);
                memcpy(qc.z_h_inv[0], P.z_h_inv[c.j].data(), 32);
                qc.m = n;
                qc.factor = 1;
                if (shuffle) {
                    uzkge_quotient_shuffle_args sc;
                    memset(&sc, 0, sizeof sc);
                    for (size_t i = 0; i < 3; i++) sc.w_sel[i] = c.dyn + 4 * n * (SLOT_W_SEL + i);
                    sc.q_ecc = c.q_ecc;
                    for (size_t i = 0; i < 12; i++) {
                        sc.pk[i] = c.pk[i];
                        sc.gen[i] = c.gen[i];
                    }
                    memcpy(sc.edwards_a, P.edwards_a.data(), 32);
                    TRY(uzkge_cuda_plonk_quotient_shuffle_fr_device(&qc, &sc, P.cbuf, st));
                } else {
                    TRY(uzkge_cuda_plonk_quotient_fr_device(&qc, P.cbuf, st));
                }
                const Limbs g_inv = FR.inverse(FR.mul(P.k1, fr_pow_u64(P.root_m, (u64)c.j)));
                TRY(uzkge_cuda_ntt_fr_device(P.cbuf, P.tcos + 4 * c.j * n, P.scratch, n, n, 1, g_inv.data(), st));
                proof->ifft_n++;
                for (size_t r = 0; r < G; r++) {
                    if (r == rank) continue;
                    Params* Q = grp->params[r];
                    CU(cudaMemcpyPeerAsync(Q->tcos + 4 * c.j * n, Q->device, P.tcos + 4 * c.j * n, P.device, n * 32, st));
                }
            }
            CU(cudaStreamSynchronize(st));
            if (!grp->bar.wait()) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: another member of the device group failed");
            TRY(uzkge_cuda_plonk_coset_combine_fr_device(P.tcos, n, P.factor, P.k1.data(), t_buf, st));
        } else {
            TRY(uzkge_cuda_ntt_fr_device(t_buf, t_buf, P.scratch, m, m, 1, P.k1_inv.data(), st));
            proof->coset_ifft_m++;
        }
    }
    size_t coefs_len = 0;
    TRY(uzkge_cuda_fr_trimmed_len_device(t_buf, m, &coefs_len, st));
    mark(2);
    // split_t_and_commit (helpers.rs:1323-1408): pieces of n + 2 coefficients, neighbouring pieces tied by one blind each
    const size_t piece = n + 2;
    {
        Limbs prev{};   // zero
        for (size_t i = 0; i < N_WIRES; i++) {
            const size_t start = i * piece;
            const size_t end = i == N_WIRES - 1 ? coefs_len : (i + 1) * piece;
            size_t take = 0;
            if (start < coefs_len) {
                const size_t hi = coefs_len < end ? coefs_len : end;
                take = hi > start ? hi - start : 0;
            }
            if (take > stride)      // the last piece of a quotient that is not a polynomial of the expected degree
                return uz::api_fail(UZKGE_ERR_SIZE, "plonk_prove: DegreeError (the quotient does not fit the SRS: unsatisfied witness?)");
            t_polys[i] = {P.poly_buf(9 + i), take};
            if (take) CU(cudaMemcpyAsync(t_polys[i].p, t_buf + 4 * start, take * 32, cudaMemcpyDeviceToDevice, st));
            const Limbs rand = next_blind();
            const Limbs neg_prev = FR.neg(prev);
            if (i != N_WIRES - 1) {
                // coefs.resize(n + 3); coefs[n + 2] += rand; coefs[0] -= prev   (helpers.rs:1351-1354)
                sparse.add(t_polys[i].p, piece, rand);
                sparse.add(t_polys[i].p, 0, neg_prev);
                t_polys[i].len = piece + 1;
            } else {
                sparse.add(t_polys[i].p, 0, neg_prev);
                if (t_polys[i].len < 1) t_polys[i].len = 1;
            }
            prev = rand;
        }
        TRY(flush_sparse());
        u64 aff[5 * 8];
        if (lagrange_all) {
            TRY(commit_coefs_lagrange(std::vector<Poly>(t_polys, t_polys + 5), aff));
        } else {
            std::vector<Msm> v;
            for (size_t i = 0; i < N_WIRES; i++) v.push_back({t_polys[i].p, t_polys[i].len});
            TRY(commit(a->srs, srs_n, v, aff, nullptr));
        }
        memcpy(proof->cm_t, aff, sizeof proof->cm_t);
        append_commitments(aff, N_WIRES);
    }
    mark(3);

    // ---- 8. zeta;  9a. the openings' values (prover.rs:217-244): one batched Horner pass, one small read
    const Limbs zeta = tr.get_challenge_field_elem();
    const Limbs zeta_omega = FR.mul(P.root, zeta);
    Limbs ev[19];
    size_t n_ev = 0;
    {
        const void* polys[19];
        size_t lens[19];
        uint32_t pt[19];
        auto push = [&](const Poly& f, uint32_t which) {
            polys[n_ev] = f.p;
            lens[n_ev] = f.len ? f.len : 1;
            pt[n_ev++] = which;
        };
        for (size_t i = 0; i < N_WIRES; i++) push(w_polys[i], 0);
        for (size_t i = 0; i < N_WIRES - 1; i++) push(P.s[i], 0);
        push(P.prk[2], 0);
        push(P.prk[3], 0);
        push(z_poly, 1);
        for (size_t i = 0; i < 3; i++) push(w_polys[i], 1);
        if (shuffle) {
            push(P.q_ecc, 0);
            for (size_t i = 0; i < 3; i++) push(w_sel_polys[i], 0);
        }
        u64 points[8];
        memcpy(points, zeta.data(), 32);
        memcpy(points + 4, zeta_omega.data(), 32);
        u64* d_vals = P.small + 16 * 12;
        if (grp) {
            // the polynomials are replicated, so the evaluations are dealt to the members (entry j to member j mod G) and shared
            // through host memory: 32 bytes each
            const void* my_polys[19];
            size_t my_lens[19], my_idx[19], mine = 0;
            uint32_t my_pt[19];
            for (size_t j = rank; j < n_ev; j += G) {
                my_polys[mine] = polys[j];
                my_lens[mine] = lens[j];
                my_pt[mine] = pt[j];
                my_idx[mine++] = j;
            }
            if (mine) {
                TRY(uzkge_cuda_poly_eval_batch_fr_device(my_polys, my_lens, my_pt, mine, points, 2, d_vals, st));
                CU(cudaMemcpyAsync(P.pinned, d_vals, mine * 32, cudaMemcpyDeviceToHost, st));
            }
            CU(cudaStreamSynchronize(st));
            for (size_t i = 0; i < mine; i++) memcpy(grp->evals.data() + 4 * my_idx[i], P.pinned + 4 * i, 32);
            if (!grp->bar.wait()) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: another member of the device group failed");
            for (size_t j = 0; j < n_ev; j++) ev[j] = load(grp->evals.data() + 4 * j);
        } else {
            TRY(uzkge_cuda_poly_eval_batch_fr_device(polys, lens, pt, n_ev, points, 2, d_vals, st));
            CU(cudaMemcpyAsync(P.pinned, d_vals, n_ev * 32, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (size_t j = 0; j < n_ev; j++) ev[j] = load(P.pinned + 4 * j);
        }
        proof->evals += (uint32_t)n_ev;
    }
    const Limbs* we = ev;               // w_polys_eval_zeta
    const Limbs* se = ev + 5;           // s_polys_eval_zeta
    const Limbs prk3 = ev[9], prk4 = ev[10], z_zo = ev[11];
    const Limbs* wo = ev + 12;          // w_polys_eval_zeta_omega
    const Limbs q_ecc_ev = shuffle ? ev[15] : Limbs{};
    const Limbs* wsl = ev + 16;         // w_sel_polys_eval_zeta
    for (size_t i = 0; i < 5; i++) tr.append_challenge(we[i]);
    for (size_t i = 0; i < 4; i++) tr.append_challenge(se[i]);
    if (shuffle)
        for (size_t i = 0; i < 3; i++) tr.append_challenge(wsl[i]);
    tr.append_challenge(prk3);
    tr.append_challenge(prk4);
    tr.append_challenge(z_zo);
    if (shuffle) tr.append_challenge(q_ecc_ev);
    for (size_t i = 0; i < 3; i++) tr.append_challenge(wo[i]);
    // 10. u (drawn to keep the transcript in step; used by the verifier only)
    (void)tr.get_challenge_field_elem();

    // ---- 9b. the linearisation polynomial r (helpers.rs:681-999)
    struct Term {
        Limbs c;
        Poly f;
    };
    std::vector<Term> terms;
    {
        const Limbs z_h_eval = fr_sub(fr_pow_u64(zeta, (u64)n), FR.one);                    // first_lagrange_poly (helpers.rs:1412-1423)
        const Limbs l1_eval = FR.mul(z_h_eval, FR.inverse(fr_sub(zeta, FR.one)));
        Limbs ap[14];
        ap[0] = FR.one;
        for (int i = 1; i < 14; i++) ap[i] = FR.mul(ap[i - 1], alpha);
        const Limbs w01 = FR.mul(we[0], we[1]), w23 = FR.mul(we[2], we[3]);
        const Limbs sel_mult[9] = {we[0], we[1], we[2], we[3], w01, w23, FR.one, FR.mul(FR.mul(w01, w23), we[4]), FR.neg(we[4])};
        for (size_t i = 0; i < N_SEL; i++) terms.push_back({sel_mult[i], P.q[i]});
        Limbs z_scalar = alpha;
        const Limbs beta_zeta = FR.mul(beta, zeta);
        for (size_t i = 0; i < N_WIRES; i++) z_scalar = FR.mul(z_scalar, FR.add(FR.add(we[i], FR.mul(P.k[i], beta_zeta)), gamma));
        z_scalar = FR.add(z_scalar, FR.mul(l1_eval, ap[2]));
        terms.push_back({z_scalar, z_poly});
        Limbs s_last = FR.mul(FR.mul(alpha, z_zo), beta);
        for (size_t i = 0; i < N_WIRES - 1; i++) s_last = FR.mul(s_last, FR.add(FR.add(we[i], FR.mul(beta, se[i])), gamma));
        terms.push_back({FR.neg(s_last), P.s[N_WIRES - 1]});
        auto bool_term = [&](const Limbs& w, const Limbs& a_) { return FR.mul(FR.mul(w, fr_sub(w, FR.one)), a_); };
        terms.push_back({FR.add(FR.add(bool_term(we[1], ap[3]), bool_term(we[2], ap[4])), bool_term(we[3], ap[5])), P.qb});
        terms.push_back({FR.mul(prk3, ap[6]), P.prk[0]});
        terms.push_back({FR.mul(prk3, ap[7]), P.prk[1]});
        if (shuffle) {
            // 6.-9. the remark-gate parts (helpers.rs:747-983), per selector combination c over x_c, y_c, dxy_c
            const Limbs one = FR.one;
            const Limbs n0 = fr_sub(one, wsl[0]), n1 = fr_sub(one, wsl[1]);
            const Limbs sel[4] = {FR.add(FR.mul(n0, n1), fr_sub(q_ecc_ev, one)), FR.mul(wsl[0], n1), FR.mul(n0, wsl[1]), FR.mul(wsl[0], wsl[1])};
            const Limbs& ed = P.edwards_a;
            for (size_t c = 0; c < 4; c++) {
                const Limbs a10 = FR.mul(ap[10], sel[c]), a11 = FR.mul(ap[11], sel[c]), a12 = FR.mul(ap[12], sel[c]), a13 = FR.mul(ap[13], sel[c]);
                terms.push_back({FR.mul(FR.mul(a10, w01), wo[0]), P.pk[8 + c]});
                terms.push_back({FR.neg(FR.mul(FR.mul(a10, wsl[2]), we[0])), P.pk[4 + c]});
                terms.push_back({FR.neg(FR.mul(a10, we[1])), P.pk[c]});
                terms.push_back({FR.neg(FR.mul(FR.mul(a11, w01), wo[1])), P.pk[8 + c]});
                terms.push_back({FR.mul(FR.mul(a11, we[0]), ed), P.pk[c]});
                terms.push_back({FR.neg(FR.mul(FR.mul(a11, wsl[2]), we[1])), P.pk[4 + c]});
                terms.push_back({FR.mul(FR.mul(a12, w23), wo[2]), P.gen[8 + c]});
                terms.push_back({FR.neg(FR.mul(FR.mul(a12, wsl[2]), we[2])), P.gen[4 + c]});
                terms.push_back({FR.neg(FR.mul(a12, we[3])), P.gen[c]});
                terms.push_back({FR.neg(FR.mul(FR.mul(a13, w23), we[4])), P.gen[8 + c]});
                terms.push_back({FR.mul(FR.mul(a13, we[2]), ed), P.gen[c]});
                terms.push_back({FR.neg(FR.mul(FR.mul(a13, wsl[2]), we[3])), P.gen[4 + c]});
            }
        }
        const Limbs zfactor = fr_pow_u64(zeta, (u64)piece);
        Limbs exponent = z_h_eval;
        for (size_t i = 0; i < N_WIRES; i++) {
            terms.push_back({FR.neg(exponent), t_polys[i]});
            exponent = FR.mul(exponent, zfactor);
        }
    }
    {
        // the same polynomial may appear in several terms (the shared zero selector, the three uses of x_c ...): merge by buffer
        std::vector<Term> merged;
        for (const Term& t : terms) {
            bool found = false;
            for (Term& mt : merged)
                if (mt.f.p == t.f.p) {
                    mt.c = FR.add(mt.c, t.c);
                    found = true;
                    break;
                }
            if (!found) merged.push_back(t);
        }
        size_t rlen = 0;
        for (const Term& t : merged)
            if (t.f.len > rlen) rlen = t.f.len;
        r_poly = {P.poly_buf(14), rlen};
        // group members on large circuits each form the coefficients [lo, hi) of r and pull the other ranges over peer memory
        const size_t lo = deal ? rank * rlen / G : 0, hi = deal ? (rank + 1) * rlen / G : rlen;
        bool first = true;
        const size_t chunk = UZKGE_LINCOMB_MAX - 1;
        for (size_t i0 = 0; i0 < merged.size() && hi > lo; i0 += chunk) {
            const void* ptrs[UZKGE_LINCOMB_MAX];
            size_t lens[UZKGE_LINCOMB_MAX];
            u64 cs[UZKGE_LINCOMB_MAX * 4];
            size_t kk = 0;
            for (size_t i = i0; i < merged.size() && i < i0 + chunk; i++) {
                const size_t len = merged[i].f.len;
                lens[kk] = len > lo ? (len < hi ? len - lo : hi - lo) : 0;
                ptrs[kk] = merged[i].f.p + 4 * (lens[kk] ? lo : 0);
                memcpy(cs + 4 * kk, merged[i].c.data(), 32);
                kk++;
            }
            if (!first) {   // accumulate onto the partial sum
                ptrs[kk] = r_poly.p + 4 * lo;
                lens[kk] = hi - lo;
                memcpy(cs + 4 * kk, FR.one.data(), 32);
                kk++;
            }
            TRY(uzkge_cuda_fr_lincomb_device(ptrs, lens, cs, kk, r_poly.p + 4 * lo, hi - lo, st));
            first = false;
        }
        if (deal) {
            CU(cudaStreamSynchronize(st));
            if (!grp->bar.wait()) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: another member of the device group failed");
            for (size_t r = 0; r < G; r++) {
                if (r == rank) continue;
                const size_t a0 = r * rlen / G, a1 = (r + 1) * rlen / G;
                Params* Q = grp->params[r];
                if (a1 > a0) CU(cudaMemcpyPeerAsync(r_poly.p + 4 * a0, P.device, Q->poly_buf(14) + 4 * a0, Q->device, (a1 - a0) * 32, st));
            }
        }
    }
    // r(zeta) is computed by the first opening's pass below: the remainder of (sum_j alpha^j p_j) / (X - zeta) is the combined value,
    // but batch_prove needs r(zeta) BEFORE its challenge, so it is one more Horner evaluation (the reference evaluates it too)
    Limbs r_eval;
    {
        u64* d_val = P.small + 16 * 12;
        TRY(uzkge_cuda_poly_horner_fr_device(r_poly.p, r_poly.len ? r_poly.len : 1, zeta.data(), nullptr, d_val, st));
        CU(cudaMemcpyAsync(P.pinned, d_val, 32, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        r_eval = load(P.pinned);
        proof->evals++;
    }
    mark(4);

    // ---- 11. the two opening proofs (batch_prove, pcs.rs:107-168); the first does not enter the transcript before the second is built
    // (prover.rs:359-381), so both quotients are committed in one batch
    u64* d_rem = P.small + 16 * 12;    // two remainders
    auto batch_prove_quotient = [&](const std::vector<Poly>& polys, const std::vector<Limbs>& evals, const Limbs& point, u64* out_q, size_t* out_len,
                                    u64* d_remainder) -> int {
        // init_pcs_batch_eval_transcript (pcs.rs:220-240)
        static const char label[] = "New PCS-Batch-Eval Protocol";
        tr.append_message(reinterpret_cast<const uint8_t*>(label), sizeof label - 1);
        const auto rbytes = uzkge::to_bytes_be(FR.modulus);
        tr.append_message(rbytes.data(), 32);
        tr.append_u64((u64)(n + 2));
        tr.append_challenge(point);
        const Limbs al = tr.get_challenge_field_elem();
        std::vector<u64> mults(4 * polys.size());
        Limbs mult = FR.one, constant{};
        size_t hlen = 0;
        std::vector<const void*> ptrs(polys.size());
        std::vector<size_t> lens(polys.size());
        for (size_t j = 0; j < polys.size(); j++) {
            memcpy(mults.data() + 4 * j, mult.data(), 32);
            constant = FR.add(constant, FR.mul(mult, evals[j]));
            mult = FR.mul(mult, al);
            ptrs[j] = polys[j].p;
            lens[j] = polys[j].len;
            if (polys[j].len > hlen) hlen = polys[j].len;
        }
        if (polys.size() > UZKGE_LINCOMB_MAX) return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: too many polynomials in one opening");
        TRY(uzkge_cuda_fr_lincomb_device(ptrs.data(), lens.data(), mults.data(), polys.size(), P.sh, hlen, st));
        const size_t idx0 = 0;
        const Limbs negc = FR.neg(constant);
        TRY(uzkge_cuda_fr_add_sparse_device(P.sh, &idx0, negc.data(), 1, st));
        TRY(uzkge_cuda_poly_horner_fr_device(P.sh, hlen, point.data(), out_q, d_remainder, st));
        *out_len = hlen - 1;
        return UZKGE_OK;
    };
    {
        std::vector<Poly> open1;
        std::vector<Limbs> evals1;
        for (size_t i = 0; i < N_WIRES; i++) {
            open1.push_back(w_polys[i]);
            evals1.push_back(we[i]);
        }
        for (size_t i = 0; i < N_WIRES - 1; i++) {
            open1.push_back(P.s[i]);
            evals1.push_back(se[i]);
        }
        open1.push_back(P.prk[2]);
        evals1.push_back(prk3);
        open1.push_back(P.prk[3]);
        evals1.push_back(prk4);
        if (shuffle) {
            open1.push_back(P.q_ecc);
            evals1.push_back(q_ecc_ev);
            for (size_t i = 0; i < 3; i++) {
                open1.push_back(w_sel_polys[i]);
                evals1.push_back(wsl[i]);
            }
        }
        open1.push_back(r_poly);
        evals1.push_back(r_eval);
        std::vector<Poly> open2 = {z_poly, w_polys[0], w_polys[1], w_polys[2]};
        std::vector<Limbs> evals2 = {z_zo, wo[0], wo[1], wo[2]};
        size_t len1 = 0, len2 = 0;
        TRY(batch_prove_quotient(open1, evals1, zeta, P.q1, &len1, d_rem));
        TRY(batch_prove_quotient(open2, evals2, zeta_omega, P.q2, &len2, d_rem + 4));
        u64 aff[16];
        if (lagrange_all) {
            TRY(commit_coefs_lagrange({Poly{P.q1, len1}, Poly{P.q2, len2}}, aff));
        } else {
            TRY(commit(a->srs, srs_n, {{P.q1, len1}, {P.q2, len2}}, aff, nullptr));
        }
        CU(cudaMemcpyAsync(P.pinned, d_rem, 64, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int i = 0; i < 8; i++)
            if (P.pinned[i]) return uz::api_fail(UZKGE_ERR_ARG, "plonk_prove: PCSProveEvalError (an opening leaves a remainder)");
        memcpy(proof->opening_witness_zeta, aff, 64);
        memcpy(proof->opening_witness_zeta_omega, aff + 8, 64);
    }
    mark(5);
    memcpy(proof->prk_3_poly_eval_zeta, prk3.data(), 32);
    memcpy(proof->prk_4_poly_eval_zeta, prk4.data(), 32);
    for (size_t i = 0; i < 5; i++) memcpy(proof->w_polys_eval_zeta[i], we[i].data(), 32);
    for (size_t i = 0; i < 3; i++) memcpy(proof->w_polys_eval_zeta_omega[i], wo[i].data(), 32);
    memcpy(proof->z_eval_zeta_omega, z_zo.data(), 32);
    for (size_t i = 0; i < 4; i++) memcpy(proof->s_polys_eval_zeta[i], se[i].data(), 32);
    if (shuffle) {
        memcpy(proof->q_ecc_poly_eval_zeta, q_ecc_ev.data(), 32);
        for (size_t i = 0; i < 3; i++) memcpy(proof->w_sel_polys_eval_zeta[i], wsl[i].data(), 32);
    }
    if (tr.state.size() == 32) memcpy(proof->transcript_state, tr.state.data(), 32);
    proof->launches = (uint32_t)(uzkge_cuda_launch_count() - launches0);
    return UZKGE_OK;
}

// the proof over a multi-device parameter handle: one worker thread per member
int prove_group(const uzkge_plonk_prove_args* a, uzkge_plonk_proof* proof, const std::vector<u64>& subs) {
    std::lock_guard<std::mutex> multi(uz::group_mutex());
    Group grp;
    const size_t G = subs.size();
    grp.G = G;
    grp.bar.members = (uint32_t)G;
    if (a->srs) {
        if (!uz::group_srs_parts_locked(a->srs, &grp.srs) || grp.srs.mode != UZKGE_MULTI_SPLIT || grp.srs.sub.size() != G)
            return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_prove: a multi-device parameter handle needs UZKGE_MULTI_SPLIT SRS handles over the same group");
    }
    if (a->lagrange_srs) {
        if (!uz::group_srs_parts_locked(a->lagrange_srs, &grp.lag) || grp.lag.mode != UZKGE_MULTI_SPLIT || grp.lag.sub.size() != G)
            return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_prove: a multi-device parameter handle needs UZKGE_MULTI_SPLIT SRS handles over the same group");
    }
    {
        std::lock_guard<std::mutex> lock(g_params_mu);
        for (u64 h : subs) {
            auto it = g_params.find(h);
            if (it == g_params.end()) return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_prove: unknown parameter handle");
            grp.params.push_back(it->second.get());
        }
    }
    grp.partial.assign(2 * G * 16 * 12, 0);
    grp.evals.assign(32 * 4, 0);
    std::vector<uzkge_plonk_proof> proofs(G);
    const int rc = uz::group_fan_out(G, [&](size_t r) {
        uzkge_plonk_prove_args mine = *a;
        mine.params = subs[r];
        const int rc_r = prove_impl(&mine, &proofs[r], &grp, r);
        if (rc_r != UZKGE_OK) grp.bar.abort();
        return rc_r;
    });
    if (rc != UZKGE_OK) return rc;
    for (size_t r = 1; r < G; r++)
        if (memcmp(proofs[r].transcript_state, proofs[0].transcript_state, 32) != 0)
            return uz::api_fail(UZKGE_ERR_INTERNAL, "plonk_prove: the members of the device group disagree on the transcript");
    *proof = proofs[0];
    return UZKGE_OK;
}

}  // namespace

extern "C" {

UZKGE_API int32_t uzkge_cuda_plonk_prove(const uzkge_plonk_prove_args* a, uzkge_plonk_proof* proof) {
    if (!a || !proof || !a->witness || !a->blinds || !a->transcript) return uz::api_fail(UZKGE_ERR_ARG, "plonk_prove: null pointer");
    if (a->params & PARAMS_MULTI) {
        std::vector<u64> subs;
        {
            std::lock_guard<std::mutex> lock(g_params_mu);
            auto it = g_multi_params.find(a->params);
            if (it == g_multi_params.end()) return uz::api_fail(UZKGE_ERR_HANDLE, "plonk_prove: unknown parameter handle");
            subs = it->second;
        }
        return prove_group(a, proof, subs);
    }
    return prove_impl(a, proof, nullptr, 0);
}

UZKGE_API int32_t uzkge_cuda_plonk_params_upload_multi(const uzkge_plonk_params_desc* d, uint64_t* params_handle) {
    if (!d || !params_handle) return uz::api_fail(UZKGE_ERR_ARG, "plonk_params_upload_multi: null pointer");
    std::vector<int> devices;
    {
        std::lock_guard<std::mutex> multi(uz::group_mutex());
        devices = uz::group_devices_locked();
    }
    if (devices.empty()) return uz::api_fail(UZKGE_ERR_NO_DEVICE, "plonk_params_upload_multi: call uzkge_cuda_init_devices first");
    const size_t G = devices.size();
    std::vector<u64> subs(G, 0);
    int rc;
    {
        std::lock_guard<std::mutex> multi(uz::group_mutex());
        rc = uz::group_fan_out(G, [&](size_t r) {
            TRY(uzkge_cuda_set_device(devices[r]));
            TRY(uzkge_cuda_plonk_params_upload(d, &subs[r]));
            Params* Pp;
            {
                std::lock_guard<std::mutex> lock(g_params_mu);
                Pp = g_params[subs[r]].get();
            }
            Params& P = *Pp;
            TRY(dev_alloc(P, P.n, &P.cbuf, false));
            TRY(dev_alloc(P, P.m, &P.tcos, false));
            TRY(build_own_cosets(P, r, G));
            return (int)UZKGE_OK;
        });
    }
    if (rc != UZKGE_OK) {
        for (u64 h : subs)
            if (h) uzkge_cuda_plonk_params_free(h);
        return rc;
    }
    std::lock_guard<std::mutex> lock(g_params_mu);
    const u64 h = PARAMS_MULTI | g_params_next++;
    g_multi_params[h] = subs;
    *params_handle = h;
    return UZKGE_OK;
}

}  // extern "C"
