// uzkge_host.hpp -- the host side above the C ABI (include/uzkge_cuda.h) in C++17, header-only: the reference's `poly_commit`
// interface for the hot path with the reference's names, argument meaning and error behaviour, so that a test written against the
// Rust types reads the same here.  (The reference is Rust; no Rust toolchain exists in this image, so the compiled-language host
// layer is C++.  The Python mirror uzkge_b200/poly_commit.py wraps the same entry points.)
//
//   UzkgeError                                   /root/reference/uzkge/src/errors.rs:6-45
//   FpPolynomial<Fr>                             /root/reference/uzkge/src/poly_commit/field_polynomial.rs:13-17, 86-90, 154-159, 198-209,
//                                                519-607
//   Radix2EvaluationDomain / MixedRadix...       ark-poly 0.4 as called from field_polynomial.rs:554-567
//   KZGCommitment, KZGCommitmentSchemeBN254      /root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:22-53, 168-204, 268-342
//   PolyComScheme::{commit, eval, prove, apply_blind_factors}
//                                                /root/reference/uzkge/src/poly_commit/pcs.rs:51-105
//
// Field elements are arkworks' in-memory form: 4 x u64 little-endian limbs, Montgomery (R = 2^256).  G1 points: affine x || y
// (8 limbs, identity = zeros) in, Jacobian X || Y || Z (12 limbs, Z = 0 identity) out.  Nothing here computes on the CPU: every
// transform, MSM, evaluation and group addition is a call into libuzkge_cuda.so, and a missing device is an error (no fallback).
#ifndef UZKGE_HOST_HPP
#define UZKGE_HOST_HPP

#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "uzkge_cuda.h"

namespace uzkge {

// ---- errors.rs: the variants the hot path can produce
enum class UzkgeError { DegreeError, CommitmentError, FFTError, ParameterError, PCSProveEvalError, Message };

struct Error : std::runtime_error {
    UzkgeError kind;
    Error(UzkgeError k, const std::string& what) : std::runtime_error(what), kind(k) {}
};

namespace detail {
// A failing C call -> the variant the Rust wrapper maps it to: size / handle / argument problems are
// ParameterError, everything else (no device, CUDA failure, out of memory) is the operation's own variant.
inline void check(int32_t rc, UzkgeError op, const char* where) {
    if (rc == UZKGE_OK) return;
    const UzkgeError kind = (rc == UZKGE_ERR_SIZE || rc == UZKGE_ERR_HANDLE || rc == UZKGE_ERR_ARG) ? UzkgeError::ParameterError : op;
    const char* msg = uzkge_cuda_last_error();
    throw Error(kind, std::string(where) + ": " + (msg ? msg : "error") + " (code " + std::to_string(rc) + ")");
}
inline bool is_pow2(size_t n) { return n > 0 && (n & (n - 1)) == 0; }
}  // namespace detail

// ---- ark_bn254::Fr as stored by arkworks
using Fr = std::array<uint64_t, 4>;
inline bool fr_is_zero(const Fr& a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
// r, little-endian limbs
inline constexpr Fr FR_MODULUS = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
// -a: r - a for a != 0 (valid in Montgomery form as well: the map x -> x R is linear)
inline Fr fr_neg(const Fr& a) {
    if (fr_is_zero(a)) return a;
    Fr out{};
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned __int128 d = (unsigned __int128)FR_MODULUS[i] - a[i] - borrow;
        out[i] = (uint64_t)d;
        borrow = (d >> 64) ? 1 : 0;
    }
    return out;
}

inline void init(int device = 0) { detail::check(uzkge_cuda_init(device), UzkgeError::Message, "uzkge_cuda_init"); }

// ---- evaluation domains: what the hot path needs of ark-poly's (size, group_gen, fft, ifft)
struct EvaluationDomain {
    size_t size_ = 0;
    Fr group_gen{};
    size_t size() const { return size_; }

    // domain.fft(&coefs): zero-pads to size(), natural order in and out
    std::vector<Fr> fft(const std::vector<Fr>& coefs) const { return transform(coefs, false, nullptr); }
    std::vector<Fr> ifft(const std::vector<Fr>& evals) const { return transform(evals, true, nullptr); }
    std::vector<Fr> transform(const std::vector<Fr>& in, bool inverse, const Fr* coset_shift) const {
        if (in.size() > size_) throw Error(UzkgeError::FFTError, "input longer than the domain");
        std::vector<Fr> buf(size_, Fr{});
        if (!in.empty()) std::memcpy(buf.data(), in.data(), in.size() * sizeof(Fr));
        detail::check(uzkge_cuda_ntt_fr(buf[0].data(), in.size(), size_, inverse ? 1 : 0, coset_shift ? coset_shift->data() : nullptr),
                      UzkgeError::FFTError, "uzkge_cuda_ntt_fr");
        return buf;
    }

  protected:
    static std::optional<EvaluationDomain> make(size_t size) {
        EvaluationDomain d;
        d.size_ = size;
        detail::check(uzkge_cuda_fr_root_of_unity(size, d.group_gen.data()), UzkgeError::FFTError, "uzkge_cuda_fr_root_of_unity");
        return d;
    }
};

struct Radix2EvaluationDomain : EvaluationDomain {
    // Radix2EvaluationDomain::new: the smallest power of two >= num_coeffs; None above 2^28 (the field's two-adicity)
    static std::optional<EvaluationDomain> new_(size_t num_coeffs) {
        size_t size = 1;
        while (size < num_coeffs) size <<= 1;
        if (size > ((size_t)1 << 28)) return std::nullopt;
        return make(size);
    }
};

struct MixedRadixEvaluationDomain : EvaluationDomain {
    // the sizes the reference asks for: 2^k or 3 * 2^k (field_polynomial.rs:561-567)
    static std::optional<EvaluationDomain> new_(size_t num_coeffs) {
        const bool ok = detail::is_pow2(num_coeffs) || (num_coeffs % 3 == 0 && detail::is_pow2(num_coeffs / 3));
        if (!ok || num_coeffs > 3 * ((size_t)1 << 28)) return std::nullopt;
        return make(num_coeffs);
    }
};

// ---- FpPolynomial: coefficient vector, low order first, trailing zeros trimmed
class FpPolynomial {
  public:
    std::vector<Fr> coefs;

    static FpPolynomial from_coefs(std::vector<Fr> c) {
        FpPolynomial p;
        p.coefs = std::move(c);
        p.trim_coefs();
        return p;
    }
    static FpPolynomial zero() { return from_coefs({Fr{}}); }

    void trim_coefs() {
        while (coefs.size() > 1 && fr_is_zero(coefs.back())) coefs.pop_back();
        if (coefs.empty()) coefs.push_back(Fr{});
    }
    const std::vector<Fr>& get_coefs_ref() const { return coefs; }
    size_t degree() const { return coefs.size() - 1; }
    bool is_zero() const { return degree() == 0 && fr_is_zero(coefs[0]); }
    bool operator==(const FpPolynomial& o) const { return coefs == o.coefs; }

    // eval (field_polynomial.rs:198-209) and div_rem by (X - z) (:519-550): the serial loops as GPU scans
    Fr eval(const Fr& point) const {
        Fr out{};
        detail::check(uzkge_cuda_poly_eval_fr(coefs[0].data(), coefs.size(), point.data(), out.data()), UzkgeError::Message,
                      "uzkge_cuda_poly_eval_fr");
        return out;
    }
    std::pair<FpPolynomial, FpPolynomial> div_rem_linear(const Fr& z) const {
        if (coefs.size() < 2) return {zero(), from_coefs(coefs)};
        std::vector<Fr> q(coefs.size() - 1);
        Fr rem{};
        detail::check(uzkge_cuda_poly_div_linear_fr(coefs[0].data(), coefs.size(), z.data(), q[0].data(), rem.data()), UzkgeError::Message,
                      "uzkge_cuda_poly_div_linear_fr");
        return {from_coefs(std::move(q)), from_coefs({rem})};
    }

    // domains (field_polynomial.rs:554-567)
    static std::optional<EvaluationDomain> evaluation_domain(size_t num_coeffs) {
        if (!detail::is_pow2(num_coeffs)) throw Error(UzkgeError::ParameterError, "evaluation_domain: not a power of two");
        return Radix2EvaluationDomain::new_(num_coeffs);
    }
    static std::optional<EvaluationDomain> quotient_evaluation_domain(size_t num_coeffs) { return MixedRadixEvaluationDomain::new_(num_coeffs); }

    // transforms (field_polynomial.rs:570-607)
    std::optional<std::vector<Fr>> fft(size_t num_coeffs) const {
        if (num_coeffs <= degree()) throw Error(UzkgeError::ParameterError, "fft: domain not larger than the degree");
        auto d = detail::is_pow2(num_coeffs) ? evaluation_domain(num_coeffs) : quotient_evaluation_domain(num_coeffs);
        if (!d) return std::nullopt;
        return fft_with_domain(*d);
    }
    std::vector<Fr> fft_with_domain(const EvaluationDomain& domain) const { return domain.transform(coefs, false, nullptr); }
    // self.mul_var(k).fft_with_domain(domain): the power scaling is fused into the transform's first read
    std::vector<Fr> coset_fft_with_domain(const EvaluationDomain& domain, const Fr& k) const { return domain.transform(coefs, false, &k); }
    static FpPolynomial ifft_with_domain(const EvaluationDomain& domain, const std::vector<Fr>& values) {
        return from_coefs(domain.transform(values, true, nullptr));
    }
    // ifft_with_domain(domain, values).mul_var(k_inv): the scaling is fused into the transform's last store
    static FpPolynomial coset_ifft_with_domain(const EvaluationDomain& domain, const std::vector<Fr>& values, const Fr& k_inv) {
        return from_coefs(domain.transform(values, true, &k_inv));
    }
};

// ---- KZGCommitment(G1Projective)
struct KZGCommitment {
    std::array<uint64_t, 12> value{};   // Jacobian X, Y, Z (Montgomery); Z = 0 is the identity

    std::array<uint64_t, 8> to_affine() const {
        std::array<uint64_t, 8> a{};
        detail::check(uzkge_cuda_g1_to_affine(value.data(), a.data()), UzkgeError::CommitmentError, "uzkge_cuda_g1_to_affine");
        return a;
    }
    bool is_identity() const { return (value[8] | value[9] | value[10] | value[11]) == 0; }
    KZGCommitment add(const KZGCommitment& o) const {
        KZGCommitment r;
        detail::check(uzkge_cuda_g1_add(value.data(), o.value.data(), r.value.data()), UzkgeError::CommitmentError, "uzkge_cuda_g1_add");
        return r;
    }
    // equality of group elements (the reference compares affine forms, kzg_poly_commitment.rs:37-53)
    bool operator==(const KZGCommitment& o) const { return to_affine() == o.to_affine(); }
};

// ---- KZG over BN254 with the G1 bases resident on the GPU (uploaded once, affine, with their window tables; the reference
// re-normalises its Vec<G1Projective> on every commit, kzg_poly_commitment.rs:287-288)
class KZGCommitmentSchemeBN254 {
  public:
    std::vector<uint64_t> public_parameter_group_1;   // n x 8 limbs, affine, identity = zeros

    explicit KZGCommitmentSchemeBN254(std::vector<uint64_t> affine_xy, uint32_t window_bits = 0) : public_parameter_group_1(std::move(affine_xy)) {
        if (public_parameter_group_1.empty() || public_parameter_group_1.size() % 8)
            throw Error(UzkgeError::ParameterError, "the SRS must hold n x 8 limbs");
        detail::check(uzkge_cuda_srs_upload(public_parameter_group_1.data(), n_points(), window_bits, &handle_), UzkgeError::CommitmentError,
                      "uzkge_cuda_srs_upload");
    }
    // KZGCommitmentScheme::new (kzg_poly_commitment.rs:183-204) with the trapdoor given explicitly: SRS[i] = tau^i G, built on the GPU
    static KZGCommitmentSchemeBN254 new_(size_t max_degree, const Fr& tau, uint32_t window_bits = 0) {
        std::vector<uint64_t> pts(8 * (max_degree + 1));
        detail::check(uzkge_cuda_srs_generate(tau.data(), max_degree + 1, pts.data()), UzkgeError::CommitmentError, "uzkge_cuda_srs_generate");
        return KZGCommitmentSchemeBN254(std::move(pts), window_bits);
    }
    // the Lagrange-basis scheme of THIS SRS over the size-n domain (what lagrange-srs-*.bin holds, gen_params/mod.rs:42-65)
    KZGCommitmentSchemeBN254 derive_lagrange(size_t n, uint32_t window_bits = 0) const {
        if (n > n_points()) throw Error(UzkgeError::ParameterError, "the monomial SRS is shorter than the domain");
        std::vector<uint64_t> pts(8 * n);
        detail::check(uzkge_cuda_srs_lagrange_from_monomial(public_parameter_group_1.data(), n, pts.data()), UzkgeError::CommitmentError,
                      "uzkge_cuda_srs_lagrange_from_monomial");
        return KZGCommitmentSchemeBN254(std::move(pts), window_bits);
    }
    KZGCommitmentSchemeBN254(const KZGCommitmentSchemeBN254&) = delete;
    KZGCommitmentSchemeBN254& operator=(const KZGCommitmentSchemeBN254&) = delete;
    KZGCommitmentSchemeBN254(KZGCommitmentSchemeBN254&& o) noexcept : public_parameter_group_1(std::move(o.public_parameter_group_1)), handle_(o.handle_) {
        o.handle_ = 0;
    }
    ~KZGCommitmentSchemeBN254() {
        if (handle_) uzkge_cuda_srs_free(handle_);
    }

    uint64_t handle() const { return handle_; }
    size_t n_points() const { return public_parameter_group_1.size() / 8; }
    size_t max_degree() const { return n_points() - 1; }

    // PolyComScheme::commit (kzg_poly_commitment.rs:278-293)
    KZGCommitment commit(const FpPolynomial& polynomial) const {
        if (polynomial.degree() + 1 > n_points()) throw Error(UzkgeError::DegreeError, "DegreeError");
        KZGCommitment c;
        detail::check(uzkge_cuda_msm_g1(handle_, 0, polynomial.coefs[0].data(), polynomial.degree() + 1, c.value.data()),
                      UzkgeError::CommitmentError, "uzkge_cuda_msm_g1");
        return c;
    }
    // the independent commitments of one prover round (plonk/prover.rs:132-192, helpers.rs:1323-1408) in one call
    std::vector<KZGCommitment> commit_batch(const std::vector<const FpPolynomial*>& polynomials) const {
        std::vector<const uint64_t*> ptrs;
        std::vector<size_t> lens;
        for (const FpPolynomial* p : polynomials) {
            if (p->degree() + 1 > n_points()) throw Error(UzkgeError::DegreeError, "DegreeError");
            ptrs.push_back(p->coefs[0].data());
            lens.push_back(p->degree() + 1);
        }
        std::vector<KZGCommitment> out(polynomials.size());
        if (out.empty()) return out;
        std::vector<uint64_t> flat(12 * out.size());
        detail::check(uzkge_cuda_msm_g1_batch(handle_, ptrs.data(), lens.data(), out.size(), flat.data()), UzkgeError::CommitmentError,
                      "uzkge_cuda_msm_g1_batch");
        for (size_t i = 0; i < out.size(); i++) std::memcpy(out[i].value.data(), flat.data() + 12 * i, 96);
        return out;
    }
    // PolyComScheme::eval (kzg_poly_commitment.rs:295-297)
    Fr eval(const FpPolynomial& poly, const Fr& point) const { return poly.eval(point); }
    // PolyComScheme::prove (kzg_poly_commitment.rs:315-342): the commitment of (P(X) - P(x)) / (X - x); the division's remainder
    // IS P(x), so the quotient of P by (X - x) is the quotient of P - P(x)
    KZGCommitment prove(const FpPolynomial& poly, const Fr& x, size_t max_degree) const {
        if (poly.degree() > max_degree) throw Error(UzkgeError::DegreeError, "DegreeError");
        return commit(poly.div_rem_linear(x).first);
    }
    // kzg_poly_commitment.rs:299-313: C + sum_i b_i (SRS[i] - SRS[zeroing_degree + i]) as two tiny MSMs over the resident bases
    KZGCommitment apply_blind_factors(const KZGCommitment& commitment, const std::vector<Fr>& blinds, size_t zeroing_degree) const {
        if (blinds.empty()) return commitment;
        if (zeroing_degree + blinds.size() > n_points()) throw Error(UzkgeError::ParameterError, "blind factors outside the SRS");
        std::vector<Fr> neg(blinds.size());
        for (size_t i = 0; i < blinds.size(); i++) neg[i] = fr_neg(blinds[i]);
        KZGCommitment lo, hi;
        detail::check(uzkge_cuda_msm_g1(handle_, 0, blinds[0].data(), blinds.size(), lo.value.data()), UzkgeError::CommitmentError,
                      "uzkge_cuda_msm_g1");
        detail::check(uzkge_cuda_msm_g1(handle_, zeroing_degree, neg[0].data(), neg.size(), hi.value.data()), UzkgeError::CommitmentError,
                      "uzkge_cuda_msm_g1");
        return commitment.add(lo).add(hi);
    }

  private:
    uint64_t handle_ = 0;
};

// ---- device-resident vectors: what keeps the prover's polynomials in HBM between the rounds.  The *_device entry
// points are called with stream = NULL; uzkge_cuda_dev_copy_in / _out are ordered with them.
class DeviceVec {
  public:
    explicit DeviceVec(size_t n_elements) : n_(n_elements) {
        detail::check(uzkge_cuda_dev_alloc(n_ * sizeof(Fr), &ptr_), UzkgeError::Message, "uzkge_cuda_dev_alloc");
    }
    explicit DeviceVec(const std::vector<Fr>& host) : DeviceVec(host.size()) { upload(host); }
    DeviceVec(const DeviceVec&) = delete;
    DeviceVec& operator=(const DeviceVec&) = delete;
    DeviceVec(DeviceVec&& o) noexcept : ptr_(o.ptr_), n_(o.n_) { o.ptr_ = nullptr; }
    ~DeviceVec() {
        if (ptr_) uzkge_cuda_dev_free(ptr_);
    }
    void* ptr() const { return ptr_; }
    size_t size() const { return n_; }
    void upload(const std::vector<Fr>& host) {
        if (host.size() > n_) throw Error(UzkgeError::ParameterError, "upload: vector longer than the buffer");
        detail::check(uzkge_cuda_dev_copy_in(ptr_, host.data(), host.size() * sizeof(Fr)), UzkgeError::Message, "uzkge_cuda_dev_copy_in");
    }
    std::vector<Fr> download(size_t count) const {
        if (count > n_) throw Error(UzkgeError::ParameterError, "download: more than the buffer holds");
        std::vector<Fr> out(count);
        detail::check(uzkge_cuda_dev_copy_out(out.data(), ptr_, count * sizeof(Fr)), UzkgeError::Message, "uzkge_cuda_dev_copy_out");
        return out;
    }

  private:
    void* ptr_ = nullptr;
    size_t n_ = 0;
};

// domain.fft / ifft (and the coset variants) of a device-resident vector: out[0..size) <- transform(in[0..len_in) zero-padded)
inline void transform_device(const EvaluationDomain& domain, const DeviceVec& in, size_t len_in, DeviceVec& out, DeviceVec& scratch, bool inverse,
                             const Fr* coset_shift = nullptr) {
    if (len_in > domain.size() || out.size() < domain.size() || scratch.size() < domain.size() || in.size() < len_in)
        throw Error(UzkgeError::FFTError, "transform_device: buffer sizes");
    detail::check(uzkge_cuda_ntt_fr_device(in.ptr(), out.ptr(), scratch.ptr(), len_in, domain.size(), inverse ? 1 : 0,
                                           coset_shift ? coset_shift->data() : nullptr, nullptr),
                  UzkgeError::FFTError, "uzkge_cuda_ntt_fr_device");
}

// PolyComScheme::commit for device-resident coefficient vectors: the independent commitments of a round in one pass
inline std::vector<KZGCommitment> commit_device(const KZGCommitmentSchemeBN254& pcs, const std::vector<const DeviceVec*>& vecs,
                                                const std::vector<size_t>& lens) {
    if (vecs.size() != lens.size()) throw Error(UzkgeError::ParameterError, "commit_device: one length per vector");
    std::vector<const void*> ptrs;
    for (size_t i = 0; i < vecs.size(); i++) {
        if (lens[i] > pcs.n_points()) throw Error(UzkgeError::DegreeError, "DegreeError");
        if (lens[i] > vecs[i]->size()) throw Error(UzkgeError::ParameterError, "commit_device: length beyond the buffer");
        ptrs.push_back(vecs[i]->ptr());
    }
    std::vector<KZGCommitment> out(vecs.size());
    if (out.empty()) return out;
    void* d_out = nullptr;
    detail::check(uzkge_cuda_dev_alloc(96 * out.size(), &d_out), UzkgeError::CommitmentError, "uzkge_cuda_dev_alloc");
    const int32_t rc = uzkge_cuda_msm_g1_batch_device(pcs.handle(), 0, ptrs.data(), lens.data(), out.size(), d_out, nullptr);
    std::vector<uint64_t> flat(12 * out.size());
    const int32_t rc2 = rc == UZKGE_OK ? uzkge_cuda_dev_copy_out(flat.data(), d_out, 96 * out.size()) : UZKGE_OK;
    uzkge_cuda_dev_free(d_out);
    detail::check(rc, UzkgeError::CommitmentError, "uzkge_cuda_msm_g1_batch_device");
    detail::check(rc2, UzkgeError::CommitmentError, "uzkge_cuda_dev_copy_out");
    for (size_t i = 0; i < out.size(); i++) std::memcpy(out[i].value.data(), flat.data() + 12 * i, 96);
    return out;
}

}  // namespace uzkge

#endif  // UZKGE_HOST_HPP
