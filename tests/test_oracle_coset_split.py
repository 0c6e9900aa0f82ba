"""CPU: the algebra the device group's quotient round rests on (csrc/prover.cu, uzkge_cuda_plonk_coset_combine_fr_device), stated
with the oracle's big-integer transforms only.  The m = 6n points k1 w_m^p of the quotient domain are the cosets g_j <w_n>,
g_j = k1 w_m^j (p = 6 i + j):
  * a polynomial of n + 3 coefficients on coset j is the size-n coset transform (shift g_j) of its first n coefficients with the
    tail folded in, X^n = g_j^n;
  * the size-n coset iFFT (shift g_j) of t's values on coset j is u_j[r] = T_r(g_j^n), where t(X) = sum_r X^r T_r(X^n);
  * a 6-point inverse DFT over the cosets gives t[r + n c] = K^-c / 6 * sum_j u_j[r] e^(j c), K = k1^n, e = (w_m^n)^-1 -- and with
    e^2 = e - 1, e^3 = -1 every e^k v is one of v, e v, e v - v and their negatives (one product per input)."""
import random

from oracle import bn254 as bn

FR = bn.FR
K1 = 0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A


def test_quotient_domain_splits_into_cosets_and_back():
    rnd = random.Random(7)
    n, f = 16, 6
    m = n * f
    w_m, w_n = bn.root_of_unity(m), bn.root_of_unity(n)
    assert pow(w_m, f, FR) == w_n
    # (1) n + 3 coefficients on coset j == size-n coset transform of the folded coefficients
    poly = [rnd.randrange(FR) for _ in range(n + 3)]
    whole = bn.coset_fft(poly, m, K1)
    for j in range(f):
        g = K1 * pow(w_m, j, FR) % FR
        gn = pow(g, n, FR)
        folded = [(poly[i] + (gn * poly[n + i] if i < 3 else 0)) % FR for i in range(n)]
        assert bn.coset_fft(folded, n, g) == whole[j::f], j
    # (2) + (3): t of 6n coefficients from its values, coset by coset
    t = [rnd.randrange(FR) for _ in range(m)]
    vals = bn.coset_fft(t, m, K1)
    u = []
    for j in range(f):
        g = K1 * pow(w_m, j, FR) % FR
        u_j = bn.coset_ifft(vals[j::f], n, bn.inv_mod(g, FR))
        gn = pow(g, n, FR)
        for r in (0, 1, n - 1):
            assert u_j[r] == sum(t[r + n * c] * pow(gn, c, FR) for c in range(f)) % FR          # u_j[r] = T_r(g_j^n)
        u.append(u_j)
    e = bn.inv_mod(pow(w_m, n, FR), FR)
    assert (e * e - e + 1) % FR == 0 and pow(e, 3, FR) == FR - 1                              # primitive 6th root: e^2 = e - 1
    K_inv, inv_f = bn.inv_mod(pow(K1, n, FR), FR), bn.inv_mod(f, FR)
    out = [0] * m
    for r in range(n):
        b = [e * u[j][r] % FR for j in range(f)]                                              # the ONE product per input
        table = lambda j: [u[j][r], b[j], (b[j] - u[j][r]) % FR, -u[j][r] % FR, -b[j] % FR, (u[j][r] - b[j]) % FR]
        for c in range(f):
            acc = sum(table(j)[(j * c) % f] for j in range(f)) % FR
            out[r + n * c] = acc * pow(K_inv, c, FR) % FR * inv_f % FR
    assert out == t
    assert bn.trim(out) == bn.trim(bn.coset_ifft(vals, m, bn.inv_mod(K1, FR)))               # what the whole-domain coset iFFT returns
