// uzkge_transcript.hpp -- the serial host part of the prover above the C ABI, in C++17 (header-only): the Fiat-Shamir transcript, the
// prover's RNG and the O(1) scalar arithmetic each round needs.  These stay on the host in the reference as well; nothing here is
// data-parallel.
//
//   Transcript                          /root/reference/uzkge/src/utils/transcript.rs:8-69, plonk/transcript.rs:8-31
//   commitment bytes                    /root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:37-53
//   ChaChaRng::from_seed, Fr::rand      rand_chacha 0.3 / ark-ff 0.4 as used by plonk/helpers.rs:147, :1349 and plonk/indexer.rs:224, :258
//   choose_ks                           /root/reference/uzkge/src/plonk/indexer.rs:211-235
//
// Keccak-256 and the ChaCha20 block come from libuzkge_host.so (uzkge_b200/csrc/hostutil.c, plain C, no CUDA).  Field elements are
// arkworks' Montgomery limbs; the Montgomery multiplication below is the host's scalar arithmetic (one product at a time).
#ifndef UZKGE_TRANSCRIPT_HPP
#define UZKGE_TRANSCRIPT_HPP

#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

extern "C" {
void uzkge_host_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);
void uzkge_host_chacha20_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]);
}

namespace uzkge {

using Limbs = std::array<uint64_t, 4>;

// ---- one prime field in Montgomery form (R = 2^256): modulus, -1 / modulus mod 2^64, R^2 mod modulus
struct MontField {
    Limbs modulus, r2, one;
    uint64_t inv;

    static bool geq(const Limbs& a, const Limbs& b) {
        for (int i = 3; i >= 0; i--)
            if (a[i] != b[i]) return a[i] > b[i];
        return true;
    }
    static Limbs sub_raw(const Limbs& a, const Limbs& b) {
        Limbs out{};
        unsigned __int128 borrow = 0;
        for (int i = 0; i < 4; i++) {
            const unsigned __int128 d = (unsigned __int128)a[i] - b[i] - borrow;
            out[i] = (uint64_t)d;
            borrow = (d >> 64) ? 1 : 0;
        }
        return out;
    }
    // a * b / R mod modulus (CIOS)
    Limbs mul(const Limbs& a, const Limbs& b) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            unsigned __int128 carry = 0;
            for (int j = 0; j < 4; j++) {
                carry += (unsigned __int128)a[j] * b[i] + t[j];
                t[j] = (uint64_t)carry;
                carry >>= 64;
            }
            carry += t[4];
            t[4] = (uint64_t)carry;
            t[5] = (uint64_t)(carry >> 64);
            const uint64_t m = t[0] * inv;
            carry = ((unsigned __int128)m * modulus[0] + t[0]) >> 64;
            for (int j = 1; j < 4; j++) {
                carry += (unsigned __int128)m * modulus[j] + t[j];
                t[j - 1] = (uint64_t)carry;
                carry >>= 64;
            }
            carry += t[4];
            t[3] = (uint64_t)carry;
            t[4] = t[5] + (uint64_t)(carry >> 64);
        }
        Limbs out = {t[0], t[1], t[2], t[3]};
        if (t[4] || geq(out, modulus)) out = sub_raw(out, modulus);
        return out;
    }
    Limbs to_mont(const Limbs& canonical) const { return mul(canonical, r2); }
    Limbs from_mont(const Limbs& m) const { return mul(m, Limbs{1, 0, 0, 0}); }
    Limbs add(const Limbs& a, const Limbs& b) const {
        Limbs out{};
        unsigned __int128 carry = 0;
        for (int i = 0; i < 4; i++) {
            carry += (unsigned __int128)a[i] + b[i];
            out[i] = (uint64_t)carry;
            carry >>= 64;
        }
        if (carry || geq(out, modulus)) out = sub_raw(out, modulus);
        return out;
    }
    Limbs neg(const Limbs& a) const { return (a[0] | a[1] | a[2] | a[3]) ? sub_raw(modulus, a) : a; }
    // a^e for a Montgomery a and a plain 256-bit exponent
    Limbs pow(const Limbs& a, const Limbs& e) const {
        Limbs acc = one;
        for (int i = 255; i >= 0; i--) {
            acc = mul(acc, acc);
            if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    Limbs inverse(const Limbs& a) const { return pow(a, sub_raw(modulus, Limbs{2, 0, 0, 0})); }   // Fermat
};

inline constexpr MontField FR = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                                 {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull},
                                 {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full},
                                 0xc2e1f593efffffffull};
inline constexpr MontField FQ = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                                 {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full},
                                 {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full},
                                 0x87d20782e4866389ull};

// canonical limbs -> 32 big-endian bytes (`into_bigint().to_bytes_be()`)
inline std::array<uint8_t, 32> to_bytes_be(const Limbs& canonical) {
    std::array<uint8_t, 32> out{};
    for (int i = 0; i < 32; i++) out[31 - i] = (uint8_t)(canonical[i >> 3] >> (8 * (i & 7)));
    return out;
}
inline Limbs from_bytes_be(const uint8_t b[32]) {
    Limbs out{};
    for (int i = 0; i < 32; i++) out[i >> 3] |= (uint64_t)b[31 - i] << (8 * (i & 7));
    return out;
}

// ---- utils/transcript.rs: Keccak-256 over a state of 32-byte slots; labels are ignored by the reference and not taken here
class Transcript {
  public:
    static constexpr size_t SLOT_SIZE = 32;
    std::vector<uint8_t> state;

    explicit Transcript(const std::vector<uint8_t>& msg) { append_message(msg.data(), msg.size()); }
    // resume from a serialised state (what the caller of uzkge_cuda_plonk_prove hands over after transcript_init_plonk)
    static Transcript from_state(const uint8_t* bytes, size_t len) {
        Transcript t("");
        t.state.assign(bytes, bytes + len);
        return t;
    }
    explicit Transcript(const char* label) { append_message(reinterpret_cast<const uint8_t*>(label), std::strlen(label)); }

    // utils/transcript.rs:20-30: short messages are left-padded to one slot, longer ones must be whole slots (the reference asserts
    // `len % SLOT_SIZE == 0`): a misaligned message fails here instead of silently diverging from the reference's transcript
    void append_message(const uint8_t* msg, size_t len) {
        if (len >= SLOT_SIZE && len % SLOT_SIZE != 0) throw std::invalid_argument("Transcript::append_message: length must be a multiple of 32");
        if (len < SLOT_SIZE) state.insert(state.end(), SLOT_SIZE - len, 0);
        state.insert(state.end(), msg, msg + len);
    }
    void append_u64(uint64_t a) {
        uint8_t buf[8];
        for (int i = 0; i < 8; i++) buf[i] = (uint8_t)(a >> (56 - 8 * i));
        append_message(buf, 8);
    }
    void append_single_byte(uint8_t b) { state.push_back(b); }
    // affine x || y (Montgomery limbs as the backend returns them; zeros = identity): canonical big-endian x || y, 64 zero bytes for
    // the identity (kzg_poly_commitment.rs:37-53)
    void append_commitment(const std::array<uint64_t, 8>& affine_mont) {
        const auto x = to_bytes_be(FQ.from_mont({affine_mont[0], affine_mont[1], affine_mont[2], affine_mont[3]}));
        const auto y = to_bytes_be(FQ.from_mont({affine_mont[4], affine_mont[5], affine_mont[6], affine_mont[7]}));
        state.insert(state.end(), x.begin(), x.end());
        state.insert(state.end(), y.begin(), y.end());
    }
    // a field element in Montgomery form
    void append_challenge(const Limbs& fr_mont) {
        const auto b = to_bytes_be(FR.from_mont(fr_mont));
        state.insert(state.end(), b.begin(), b.end());
    }
    // hash -> integer (big-endian) mod r -> state := that value; returned in Montgomery form, ready for the device entry points
    Limbs get_challenge_field_elem() {
        uint8_t digest[32];
        uzkge_host_keccak256(state.data(), state.size(), digest);
        Limbs v = from_bytes_be(digest);
        while (MontField::geq(v, FR.modulus)) v = MontField::sub_raw(v, FR.modulus);      // 2^256 < 6 r: at most 5 subtractions
        const auto b = to_bytes_be(v);
        state.assign(b.begin(), b.end());
        return FR.to_mont(v);
    }
};

// ---- rand_chacha::ChaChaRng (ChaCha20, 64-bit block counter, stream 0) and arkworks' Fr::rand on top of it
class ChaChaRng {
  public:
    explicit ChaChaRng(const std::array<uint8_t, 32>& seed) {
        for (int i = 0; i < 8; i++) key_[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    }
    static ChaChaRng from_seed(const std::array<uint8_t, 32>& seed) { return ChaChaRng(seed); }
    uint32_t next_u32() {
        if (pos_ >= 16) {
            uzkge_host_chacha20_block(key_, counter_++, 0, buf_);
            pos_ = 0;
        }
        return buf_[pos_++];
    }
    uint64_t next_u64() {
        const uint64_t lo = next_u32();
        return lo | ((uint64_t)next_u32() << 32);
    }

  private:
    uint32_t key_[8] = {}, buf_[16] = {};
    uint64_t counter_ = 0;
    int pos_ = 16;
};

// Fr::rand (ark-ff 0.4): 4 x next_u64 as the RAW Montgomery limbs, top 2 bits cleared, redrawn while >= r.  Returns Montgomery form.
inline Limbs fr_rand(ChaChaRng& prng) {
    for (;;) {
        Limbs v = {prng.next_u64(), prng.next_u64(), prng.next_u64(), prng.next_u64()};
        v[3] &= (1ull << 62) - 1;
        if (!MontField::geq(v, FR.modulus)) return v;
    }
}

// plonk/indexer.rs:211-235: k[0] = 1, then distinct quadratic non-residues drawn from `prng` (Montgomery form)
inline std::vector<Limbs> choose_ks(ChaChaRng& prng, size_t n_wires_per_gate) {
    std::vector<Limbs> k = {FR.one};
    Limbs half = MontField::sub_raw(FR.modulus, Limbs{1, 0, 0, 0});      // (r - 1) / 2
    for (int i = 0; i < 4; i++) half[i] = (half[i] >> 1) | (i < 3 ? half[i + 1] << 63 : 0);
    while (k.size() < n_wires_per_gate) {
        const Limbs ki = fr_rand(prng);
        if ((ki[0] | ki[1] | ki[2] | ki[3]) == 0) continue;
        bool seen = false;
        for (const Limbs& x : k) seen = seen || x == ki;
        if (!seen && FR.pow(ki, half) != FR.one) k.push_back(ki);
    }
    return k;
}

}  // namespace uzkge

#endif  // UZKGE_TRANSCRIPT_HPP
