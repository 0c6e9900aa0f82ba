// The reference's unit tests for the hot-path boundary, restated against the C++ host layer (include/uzkge_host.hpp):
//   test_fft / check_fft                      /root/reference/uzkge/src/poly_commit/field_polynomial.rs:632-719
//   test_commit, test_homomorphic_poly_com_elem  /root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:483-548
//   test_pcs_eval                              /root/reference/uzkge/src/poly_commit/pcs.rs:260-290
// Expected values come from the CPU oracle (oracle/oracle.c) -- this is a TEST program; the product never links the oracle.
//   host_api_test nodevice   (no GPU visible: every device call must fail loudly with the mapped UzkgeError; host logic is checked)
//   host_api_test gpu        (parity with the oracle through the C++ layer)
#include <cstdio>
#include <cstdlib>
#include <string>

#include "uzkge_host.hpp"
#include "uzkge_transcript.hpp"

extern "C" {
void oracle_init(void);
void oracle_fr_to_mont(const uint64_t* a, uint64_t* o, size_t n);
int oracle_msm_g1(const uint64_t* bases, const uint64_t* scalars, size_t n, uint64_t* out_jac);
void oracle_g1_mul(const uint64_t* base_aff, const uint64_t* scalar_mont, uint64_t* out_jac);
void oracle_g1_add_jac(const uint64_t* a, const uint64_t* b, uint64_t* out_jac);
void oracle_g1_to_affine(const uint64_t* in_jac, uint64_t* out_aff);
int oracle_ntt_fr(uint64_t* data, size_t len_in, size_t n, int inverse, const uint64_t* coset);
void oracle_fr_eval(const uint64_t* coefs, size_t n, const uint64_t* x_mont, uint64_t* out_mont);
void oracle_fr_inv(const uint64_t* a_mont, uint64_t* out_mont);
void oracle_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n);
void oracle_fq_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n);
void oracle_fr_from_mont(const uint64_t* a, uint64_t* o, size_t n);
void oracle_fq_to_mont(const uint64_t* a, uint64_t* o, size_t n);
}

using namespace uzkge;

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            failures++;                                                    \
        }                                                                  \
    } while (0)

template <class F>
static bool throws(UzkgeError kind, F&& f) {
    try {
        f();
    } catch (const Error& e) {
        return e.kind == kind;
    }
    return false;
}

static Fr mont(uint64_t v) {   // a small integer in Montgomery form
    Fr c = {v, 0, 0, 0}, m{};
    oracle_fr_to_mont(c.data(), m.data(), 1);
    return m;
}
static std::vector<Fr> pseudo_random(size_t n, uint64_t seed) {   // Montgomery images of 62-bit values: valid field elements
    std::vector<Fr> out(n);
    uint64_t s = seed;
    for (auto& x : out) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        x = mont(s >> 2);
    }
    return out;
}
static std::array<uint64_t, 8> oracle_affine(const std::array<uint64_t, 12>& jac) {
    std::array<uint64_t, 8> a{};
    oracle_g1_to_affine(jac.data(), a.data());
    return a;
}

static void host_logic() {
    // from_coefs trims trailing zeros, keeps one coefficient (field_polynomial.rs:154-159)
    auto p = FpPolynomial::from_coefs({mont(1), mont(2), Fr{}, Fr{}});
    CHECK(p.degree() == 1 && p.coefs.size() == 2);
    CHECK(FpPolynomial::from_coefs({}).is_zero() && FpPolynomial::zero().degree() == 0);
    CHECK(FpPolynomial::from_coefs({Fr{}, Fr{}}).is_zero());
    // -x + x = 0 mod r on the limbs; -0 = 0
    Fr one = mont(1), neg = fr_neg(one);
    unsigned __int128 carry = 0;
    Fr sum{};
    for (int i = 0; i < 4; i++) {
        carry += (unsigned __int128)one[i] + neg[i];
        sum[i] = (uint64_t)carry;
        carry >>= 64;
    }
    CHECK(sum == FR_MODULUS && fr_is_zero(fr_neg(Fr{})));
    CHECK(throws(UzkgeError::ParameterError, [] { FpPolynomial::evaluation_domain(12); }));
}

static int run_nodevice() {
    host_logic();
    CHECK(uzkge_cuda_device_count() <= 0);
    // no CPU fallback: every entry point of the layer reports the operation's error
    CHECK(throws(UzkgeError::FFTError, [] { Radix2EvaluationDomain::new_(8).value().fft({mont(1)}); }) ||
          throws(UzkgeError::FFTError, [] { Radix2EvaluationDomain::new_(8); }));
    CHECK(throws(UzkgeError::CommitmentError, [] { KZGCommitmentSchemeBN254::new_(4, mont(7)); }));
    CHECK(throws(UzkgeError::CommitmentError, [] { KZGCommitmentSchemeBN254(std::vector<uint64_t>(16, 1)); }));
    CHECK(throws(UzkgeError::ParameterError, [] { KZGCommitmentSchemeBN254(std::vector<uint64_t>(7, 1)); }));
    CHECK(throws(UzkgeError::Message, [] { FpPolynomial::from_coefs({mont(1), mont(2)}).eval(mont(3)); }));
    CHECK(throws(UzkgeError::CommitmentError, [] { KZGCommitment{}.add(KZGCommitment{}); }));
    CHECK(throws(UzkgeError::Message, [] { DeviceVec v(4); }));
    CHECK(MixedRadixEvaluationDomain::new_(10) == std::nullopt ? true : false);
    return failures;
}

static void test_fft() {
    Fr k = mont(0x1234567);   // a coset shift
    Fr k_inv{};
    oracle_fr_inv(k.data(), k_inv.data());
    for (size_t n : {1, 2, 8, 64, 1024, 12, 96, 3072}) {
        const bool radix2 = (n & (n - 1)) == 0;
        auto domain = radix2 ? FpPolynomial::evaluation_domain(n) : FpPolynomial::quotient_evaluation_domain(n);
        CHECK(domain.has_value() && domain->size() == n);
        for (size_t len : {n, n / 2 + 1}) {
            auto poly = FpPolynomial::from_coefs(pseudo_random(len, 17 * n + len));
            // check_fft (field_polynomial.rs:632-646): the transform against the oracle's, natural order, zero-padded input
            std::vector<Fr> want(n, Fr{});
            std::memcpy(want.data(), poly.coefs.data(), poly.coefs.size() * sizeof(Fr));
            CHECK(oracle_ntt_fr(want[0].data(), poly.coefs.size(), n, 0, nullptr) == 0);
            auto evals = poly.fft_with_domain(*domain);
            CHECK(evals == want);
            CHECK(poly.fft(n).value() == want);
            // test_fft: ifft(fft(p)) == p, trimmed; and the coset pair with the scaling fused
            CHECK(FpPolynomial::ifft_with_domain(*domain, evals) == poly);
            std::vector<Fr> cwant(n, Fr{});
            std::memcpy(cwant.data(), poly.coefs.data(), poly.coefs.size() * sizeof(Fr));
            CHECK(oracle_ntt_fr(cwant[0].data(), poly.coefs.size(), n, 0, k.data()) == 0);
            auto cevals = poly.coset_fft_with_domain(*domain, k);
            CHECK(cevals == cwant);
            CHECK(FpPolynomial::coset_ifft_with_domain(*domain, cevals, k_inv) == poly);
        }
    }
    CHECK(throws(UzkgeError::FFTError, [] { Radix2EvaluationDomain::new_(4).value().fft(pseudo_random(5, 1)); }));
    CHECK(!MixedRadixEvaluationDomain::new_(10).has_value());
}

static void test_commit_and_pcs() {
    const size_t max_degree = 40;
    Fr tau = mont(0xC0FFEE1234ull);
    auto pcs = KZGCommitmentSchemeBN254::new_(max_degree, tau);
    CHECK(pcs.max_degree() == max_degree);
    // test_commit: the commitment is the MSM of the coefficients over the SRS
    auto p = FpPolynomial::from_coefs(pseudo_random(max_degree + 1, 5));
    auto cm = pcs.commit(p);
    std::array<uint64_t, 12> want{};
    CHECK(oracle_msm_g1(pcs.public_parameter_group_1.data(), p.coefs[0].data(), p.coefs.size(), want.data()) == 0);
    CHECK(cm.to_affine() == oracle_affine(want));
    CHECK(throws(UzkgeError::DegreeError, [&] { pcs.commit(FpPolynomial::from_coefs(pseudo_random(max_degree + 2, 6))); }));
    // test_homomorphic_poly_com_elem on polynomials with disjoint supports: com(p) + com(q) == com(p + q)
    auto a = pseudo_random(3, 9);
    auto lo = FpPolynomial::from_coefs({a[0], a[1]});
    auto hi = FpPolynomial::from_coefs({Fr{}, Fr{}, a[2]});
    auto both = FpPolynomial::from_coefs({a[0], a[1], a[2]});
    CHECK(pcs.commit(lo).add(pcs.commit(hi)) == pcs.commit(both));
    auto batch = pcs.commit_batch({&lo, &hi, &both, &p});
    CHECK(batch.size() == 4 && batch[0] == pcs.commit(lo) && batch[2] == pcs.commit(both) && batch[3] == cm);
    // test_pcs_eval: eval, and prove = the commitment of (P - P(x)) / (X - x)
    Fr x = mont(0x55AA55AA);
    Fr ev = pcs.eval(p, x), ev_want{};
    oracle_fr_eval(p.coefs[0].data(), p.coefs.size(), x.data(), ev_want.data());
    CHECK(ev == ev_want);
    auto qr = p.div_rem_linear(x);
    CHECK(qr.second.coefs[0] == ev && qr.first.degree() == p.degree() - 1);
    CHECK(pcs.prove(p, x, max_degree) == pcs.commit(qr.first));
    CHECK(throws(UzkgeError::DegreeError, [&] { pcs.prove(p, x, 3); }));
    // apply_blind_factors: C + b0 (SRS[0] - SRS[n]) + b1 (SRS[1] - SRS[n + 1])
    const size_t n = 16;
    auto blinds = pseudo_random(2, 11);
    auto blinded = pcs.apply_blind_factors(cm, blinds, n);
    std::array<uint64_t, 12> acc = cm.value, t{};
    for (size_t i = 0; i < 2; i++) {
        Fr nb = fr_neg(blinds[i]);
        oracle_g1_mul(pcs.public_parameter_group_1.data() + 8 * i, blinds[i].data(), t.data());
        oracle_g1_add_jac(acc.data(), t.data(), acc.data());
        oracle_g1_mul(pcs.public_parameter_group_1.data() + 8 * (n + i), nb.data(), t.data());
        oracle_g1_add_jac(acc.data(), t.data(), acc.data());
    }
    CHECK(blinded.to_affine() == oracle_affine(acc));
    CHECK(throws(UzkgeError::ParameterError, [&] { pcs.apply_blind_factors(cm, blinds, max_degree); }));
    // the Lagrange SRS of the same trapdoor commits evaluation vectors to the same group element (prover_with_lagrange)
    auto lag = pcs.derive_lagrange(n);
    auto small = FpPolynomial::from_coefs(pseudo_random(n, 13));
    auto evals = small.fft_with_domain(*FpPolynomial::evaluation_domain(n));
    CHECK(lag.commit(FpPolynomial{evals}) == pcs.commit(small));
}

// the same results with the polynomials resident in HBM (the *_device rows of the ABI): upload once, transform and commit there
static void test_device_resident() {
    const size_t n = 1024, m = 6 * n;
    auto pcs = KZGCommitmentSchemeBN254::new_(n + 2, mont(0xABCDEF0123ull));
    auto p = FpPolynomial::from_coefs(pseudo_random(n, 21)), q = FpPolynomial::from_coefs(pseudo_random(n + 3, 22));
    DeviceVec dp(p.coefs), dq(q.coefs), evals(m), back(m), scratch(m);
    CHECK(dp.download(n) == p.coefs);
    auto cms = commit_device(pcs, {&dp, &dq}, {p.coefs.size(), q.coefs.size()});
    CHECK(cms.size() == 2 && cms[0] == pcs.commit(p) && cms[1] == pcs.commit(q));
    Fr k = mont(0x77123), k_inv{};
    oracle_fr_inv(k.data(), k_inv.data());
    auto dom_n = *FpPolynomial::evaluation_domain(n);
    auto dom_m = *FpPolynomial::quotient_evaluation_domain(m);
    transform_device(dom_n, dp, n, evals, scratch, false);
    CHECK(evals.download(n) == p.fft_with_domain(dom_n));
    transform_device(dom_m, dq, n + 3, evals, scratch, false, &k);                // coset_fft of n + 3 coefficients on the 6n domain
    CHECK(evals.download(m) == q.coset_fft_with_domain(dom_m, k));
    transform_device(dom_m, evals, m, back, scratch, true, &k_inv);               // and back
    auto coefs = back.download(m);
    CHECK(FpPolynomial::from_coefs(coefs) == q);
    CHECK(throws(UzkgeError::DegreeError, [&] { commit_device(pcs, {&evals}, {m}); }));
    CHECK(throws(UzkgeError::FFTError, [&] { transform_device(dom_m, dp, n, dp, scratch, false); }));
}

static Limbs parse_hex(const char* h) {           // "0x..." canonical integer -> limbs
    Limbs out{};
    std::string t(h);
    if (t.rfind("0x", 0) == 0) t = t.substr(2);
    int nib = 0;
    for (auto it = t.rbegin(); it != t.rend(); ++it, ++nib) {
        const char c = *it;
        const uint64_t v = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
        out[nib >> 4] |= v << (4 * (nib & 15));
    }
    return out;
}
static void print_hex(const char* name, const Limbs& canonical) {
    std::printf("%s 0x%016llx%016llx%016llx%016llx\n", name, (unsigned long long)canonical[3], (unsigned long long)canonical[2],
                (unsigned long long)canonical[1], (unsigned long long)canonical[0]);
}

// the serial host part (include/uzkge_transcript.hpp): needs no device.  argv: the golden k[1..4] of the reference's verifier keys.
static int run_serial(int argc, char** argv) {
    // Keccak-256 of the empty string
    uint8_t d[32];
    uzkge_host_keccak256(nullptr, 0, d);
    CHECK(from_bytes_be(d) == parse_hex("0xc5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"));
    // the host's Montgomery arithmetic against the oracle's, both fields
    auto xs = pseudo_random(64, 3), ys = pseudo_random(64, 4);
    for (size_t i = 0; i < xs.size(); i++) {
        Limbs want{}, back{};
        oracle_fr_mul(xs[i].data(), ys[i].data(), want.data(), 1);
        CHECK(FR.mul(xs[i], ys[i]) == want);
        oracle_fq_mul(xs[i].data(), ys[i].data(), want.data(), 1);       // any limbs below p are valid Fq elements too
        CHECK(FQ.mul(xs[i], ys[i]) == want);
        oracle_fr_from_mont(xs[i].data(), back.data(), 1);
        CHECK(FR.from_mont(xs[i]) == back && FR.to_mont(back) == xs[i]);
        CHECK(FR.mul(xs[i], FR.inverse(xs[i])) == FR.one);
        CHECK(FR.add(xs[i], FR.neg(xs[i])) == Limbs{} && FR.neg(xs[i]) == fr_neg(xs[i]));
    }
    CHECK(FR.from_mont(FR.one) == Limbs({1, 0, 0, 0}) && FQ.from_mont(FQ.one) == Limbs({1, 0, 0, 0}));
    CHECK(FR.modulus == FR_MODULUS);
    // choose_ks with ChaChaRng::from_seed([0; 32]) (plonk/indexer.rs:258): the k written into the reference's verifier keys
    ChaChaRng prng(std::array<uint8_t, 32>{});
    auto k = choose_ks(prng, 5);
    CHECK(k.size() == 5 && k[0] == FR.one && argc >= 6);
    for (int i = 1; i < 5 && i + 1 < argc; i++) CHECK(FR.from_mont(k[i]) == parse_hex(argv[i + 1]));
    // a fixed transcript script; the caller compares the two challenges with the Python mirror's
    Transcript tr("Plonk shuffle Proof");
    tr.append_u64(52);
    tr.append_challenge(k[1]);
    std::array<uint64_t, 8> g{};
    const Limbs gx = FQ.to_mont({1, 0, 0, 0}), gy = FQ.to_mont({2, 0, 0, 0});
    std::memcpy(g.data(), gx.data(), 32);
    std::memcpy(g.data() + 4, gy.data(), 32);
    tr.append_commitment(g);
    tr.append_commitment(std::array<uint64_t, 8>{});      // the identity: 64 zero bytes
    const Limbs c1 = tr.get_challenge_field_elem();
    tr.append_single_byte(0x01);
    tr.append_challenge(FR.mul(c1, k[2]));
    const Limbs c2 = tr.get_challenge_field_elem();
    print_hex("challenge1", FR.from_mont(c1));
    print_hex("challenge2", FR.from_mont(c2));
    // utils/transcript.rs:20-30 asserts whole slots for messages of 32 bytes and more: the C++ twin must refuse them too
    bool refused = false;
    try {
        const uint8_t junk[40] = {};
        tr.append_message(junk, sizeof junk);
    } catch (const std::invalid_argument&) {
        refused = true;
    }
    CHECK(refused);
    return failures;
}

// ---- a compiled host calling the prover behind the C ABI (what the patched Rust body of prover_with_lagrange does): the job file
// (written by tests/test_gpu_zz_host_cpp.py from the zshuffle circuit) holds tagged records  u64 tag | u64 bytes | payload.
// Prints the proof in PlonkProof::to_bytes_be order (plonk/indexer.rs:538-590) as hex; the Python test compares it with plonk.py's.
namespace job {
enum Tag : uint64_t { SIZES = 1, WIRING, PERM, K, Q, S, QB, PRK, ANEMOI, PUB_ROWS, PUB_WIT, Q_ECC, GEN, PK, EDWARDS, WITNESS, W_SEL, BLINDS, TRANSCRIPT,
                      SRS, LAGRANGE, FLAGS };
struct Rec {
    uint64_t tag;
    std::vector<uint8_t> data;
};
static std::vector<Rec> load(const char* path) {
    std::vector<Rec> out;
    FILE* f = std::fopen(path, "rb");
    if (!f) return out;
    uint64_t hdr[2];
    while (std::fread(hdr, 8, 2, f) == 2) {
        Rec r{hdr[0], std::vector<uint8_t>(hdr[1])};
        if (hdr[1] && std::fread(r.data.data(), 1, hdr[1], f) != hdr[1]) break;
        out.push_back(std::move(r));
    }
    std::fclose(f);
    return out;
}
}  // namespace job

static void put_point(std::string& hex, const uint64_t aff[8]) {
    static const char* d = "0123456789abcdef";
    for (int c = 0; c < 2; c++) {
        const auto b = to_bytes_be(FQ.from_mont({aff[4 * c], aff[4 * c + 1], aff[4 * c + 2], aff[4 * c + 3]}));
        for (uint8_t v : b) {
            hex.push_back(d[v >> 4]);
            hex.push_back(d[v & 15]);
        }
    }
}
static void put_scalar(std::string& hex, const uint64_t m[4]) {
    static const char* d = "0123456789abcdef";
    const auto b = to_bytes_be(FR.from_mont({m[0], m[1], m[2], m[3]}));
    for (uint8_t v : b) {
        hex.push_back(d[v >> 4]);
        hex.push_back(d[v & 15]);
    }
}

static int run_prove(const char* path) {
    const auto recs = job::load(path);
    if (recs.empty()) {
        std::printf("FAIL: cannot read %s\n", path);
        return 1;
    }
    uzkge_plonk_params_desc d;
    std::memset(&d, 0, sizeof d);
    uzkge_plonk_prove_args a;
    std::memset(&a, 0, sizeof a);
    const uint64_t *srs = nullptr, *lag = nullptr;
    size_t srs_n = 0, lag_n = 0, nq = 0, ns = 0, nprk = 0, ngen = 0, npk = 0, nsel = 0;
    for (const auto& r : recs) {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(r.data.data());
        const size_t elems = r.data.size() / 32;
        switch (r.tag) {
            case job::SIZES: d.n = w[0]; d.m = w[1]; d.num_vars = w[2]; d.shuffle = (int32_t)w[3]; break;
            case job::WIRING: d.wiring = reinterpret_cast<const uint32_t*>(w); break;
            case job::PERM: d.permutation = w; break;
            case job::K: std::memcpy(d.k, w, sizeof d.k); break;
            case job::Q: d.q_polys[nq] = elems ? w : nullptr; d.q_len[nq++] = elems; break;
            case job::S: d.s_polys[ns] = elems ? w : nullptr; d.s_len[ns++] = elems; break;
            case job::QB: d.qb_poly = elems ? w : nullptr; d.qb_len = elems; break;
            case job::PRK: d.q_prk_polys[nprk] = elems ? w : nullptr; d.q_prk_len[nprk++] = elems; break;
            case job::ANEMOI: std::memcpy(d.anemoi_generator, w, 32); std::memcpy(d.anemoi_generator_inv, w + 4, 32); break;
            case job::PUB_ROWS: d.public_vars_constraint_indices = w; d.n_public = r.data.size() / 8; break;
            case job::PUB_WIT: d.public_vars_witness_indices = w; break;
            case job::Q_ECC: d.q_ecc_poly = elems ? w : nullptr; d.q_ecc_len = elems; break;
            case job::GEN: d.q_shuffle_generator_polys[ngen] = elems ? w : nullptr; d.gen_len[ngen++] = elems; break;
            case job::PK: d.q_shuffle_public_key_polys[npk] = elems ? w : nullptr; d.pk_len[npk++] = elems; break;
            case job::EDWARDS: std::memcpy(d.edwards_a, w, 32); break;
            case job::WITNESS: a.witness = w; break;
            case job::W_SEL: a.w_sel_evals[nsel++] = w; break;
            case job::BLINDS: a.blinds = w; a.n_blinds = elems; break;
            case job::TRANSCRIPT: a.transcript = r.data.data(); a.transcript_len = r.data.size(); break;
            case job::SRS: srs = w; srs_n = r.data.size() / 64; break;
            case job::LAGRANGE: lag = w; lag_n = r.data.size() / 64; break;
            case job::FLAGS: a.lagrange_all = (int32_t)w[0]; break;
            default: break;
        }
    }
    auto ok = [](int32_t rc, const char* what) {
        if (rc != UZKGE_OK) std::printf("FAIL %s: %d %s\n", what, rc, uzkge_cuda_last_error());
        return rc == UZKGE_OK;
    };
    if (!ok(uzkge_cuda_init(0), "init")) return 1;
    if (!ok(uzkge_cuda_srs_upload(srs, srs_n, 0, &a.srs), "srs_upload")) return 1;
    if (lag && !ok(uzkge_cuda_srs_upload_lagrange_commit(lag, lag_n, srs, srs_n, 0, &a.lagrange_srs), "srs_upload_lagrange_commit")) return 1;
    if (!ok(uzkge_cuda_plonk_params_upload(&d, &a.params), "plonk_params_upload")) return 1;
    uzkge_plonk_proof pr;
    if (!ok(uzkge_cuda_plonk_prove(&a, &pr), "plonk_prove")) return 1;
    std::string hex;
    for (int i = 0; i < 5; i++) put_point(hex, pr.cm_w[i]);
    if (d.shuffle)
        for (int i = 0; i < 3; i++) put_point(hex, pr.cm_w_sel[i]);
    for (int i = 0; i < 5; i++) put_point(hex, pr.cm_t[i]);
    put_point(hex, pr.cm_z);
    put_scalar(hex, pr.prk_3_poly_eval_zeta);
    put_scalar(hex, pr.prk_4_poly_eval_zeta);
    for (int i = 0; i < 5; i++) put_scalar(hex, pr.w_polys_eval_zeta[i]);
    for (int i = 0; i < 3; i++) put_scalar(hex, pr.w_polys_eval_zeta_omega[i]);
    put_scalar(hex, pr.z_eval_zeta_omega);
    for (int i = 0; i < 4; i++) put_scalar(hex, pr.s_polys_eval_zeta[i]);
    if (d.shuffle) {
        put_scalar(hex, pr.q_ecc_poly_eval_zeta);
        for (int i = 0; i < 3; i++) put_scalar(hex, pr.w_sel_polys_eval_zeta[i]);
    }
    put_point(hex, pr.opening_witness_zeta);
    put_point(hex, pr.opening_witness_zeta_omega);
    std::printf("proof %s\nmsm %u launches %u\nPASS prove\n", hex.c_str(), pr.msm, pr.launches);
    uzkge_cuda_plonk_params_free(a.params);
    uzkge_cuda_srs_free(a.srs);
    if (a.lagrange_srs) uzkge_cuda_srs_free(a.lagrange_srs);
    return 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "gpu";
    oracle_init();
    if (mode == "prove") return run_prove(argc > 2 ? argv[2] : "");
    if (mode == "serial") {
        const int f = run_serial(argc, argv);
        std::printf(f ? "FAILED (%d)\n" : "PASS serial\n", f);
        return f ? 1 : 0;
    }
    if (mode == "nodevice") {
        const int f = run_nodevice();
        std::printf(f ? "FAILED (%d)\n" : "PASS nodevice\n", f);
        return f ? 1 : 0;
    }
    try {
        init(0);
        host_logic();
        test_fft();
        test_commit_and_pcs();
        test_device_resident();
    } catch (const Error& e) {
        std::printf("FAIL: uncaught uzkge::Error: %s\n", e.what());
        return 1;
    }
    std::printf(failures ? "FAILED (%d)\n" : "PASS gpu\n", failures);
    return failures ? 1 : 0;
}
