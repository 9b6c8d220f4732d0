"""CPU ORACLE (test infrastructure, NOT a product path) -- exact big-integer restatement of the
BLS12-377 arithmetic that snarkVM 0.14.5 executes behind the Aleo SDK's proving calls.

PARITY UNPINNED at the MSM / FFT boundary: the reference (/root/reference, demox-labs/aleo =
Aleo SDK 0.5.1) holds no golden vector for `VariableBase::msm` or `EvaluationDomain::fft*`;
the algorithm lives in the crates.io dependencies snarkvm-algorithms / -curves / -fields /
-utilities, all pinned "=0.14.5" (reference Cargo.toml:28-53, Cargo.lock:2200-2203, 2637-2640,
2652-2655, 2860-2863) whose sources are absent from this environment.  What the reference DOES
pin and this file is checked against (tests/test_oracle.py):
  * the real proof string in wasm/src/programs/transaction.rs:100 -> 13 compressed G1 points that
    must decompress onto y^2 = x^3 + 1 and lie in the r-torsion, Fr evaluations < r;
  * the curve/field parameters (re-derived from the BLS12 family parameter x).
Everything else is anchored on mathematics: MSM and NTT outputs are exact elements of a finite
group/field, so an independent exact implementation is a valid bit-exactness oracle.

Conventions restated (snarkVM 0.14.5):
  * Fp256/Fp384 hold Montgomery residues a*R mod m, little-endian u64 limbs, R = 2^256 / 2^384.
  * MSM scalars are canonical (non-Montgomery) BigInteger256, little-endian u64 limbs, < r.
  * G1Affine = {x: Fq, y: Fq, infinity: bool} -> 104-byte stride (96-byte payload + flag + pad).
  * G1Projective = Jacobian {x, y, z}, identity has z = 0 (snarkVM uses (0, 1, 0)).
  * EvaluationDomain: omega = TWO_ADIC_ROOT_OF_UNITY^(2^(47-k)); fft is natural in / natural out;
    ifft uses omega^-1 then scales by n^-1; coset_fft pre-multiplies c_i by g^i (g = 22);
    coset_ifft = ifft then multiply by g^-i; fft_in_place zero-pads (resize) to the domain size.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import struct

# ----------------------------------------------------------------------------------------------
# Parameters (BLS12 family, x = 0x8508c00000000001)
# ----------------------------------------------------------------------------------------------
X_PARAM = 0x8508C00000000001
R_MOD = X_PARAM**4 - X_PARAM**2 + 1                       # scalar field Fr (253 bits)
P_MOD = (X_PARAM - 1) ** 2 * R_MOD // 3 + X_PARAM          # base field Fq (377 bits)
assert R_MOD == 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
assert P_MOD == 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001

FR_LIMBS64, FQ_LIMBS64 = 4, 6
FR_R = (1 << 256) % R_MOD          # Montgomery R for Fr
FQ_R = (1 << 384) % P_MOD          # Montgomery R for Fq
FR_RINV = pow(1 << 256, -1, R_MOD)
FQ_RINV = pow(1 << 384, -1, P_MOD)
FR_INV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)   # 0x0a117fffffffffff
FQ_INV64 = (-pow(P_MOD, -1, 1 << 64)) % (1 << 64)   # 0x8508bfffffffffff
FR_INV32 = FR_INV64 & 0xFFFFFFFF
FQ_INV32 = FQ_INV64 & 0xFFFFFFFF

FR_TWO_ADICITY = 47
FR_GENERATOR = 22                  # multiplicative generator == coset shift used by snarkVM
FR_T = (R_MOD - 1) >> FR_TWO_ADICITY
FR_TWO_ADIC_ROOT = pow(FR_GENERATOR, FR_T, R_MOD)

COEFF_B = 1                        # E: y^2 = x^3 + 1
G1_COFACTOR = (X_PARAM - 1) ** 2 // 3
G1_GEN = (
    81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
    241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030,
)


# ----------------------------------------------------------------------------------------------
# Limb / byte helpers
# ----------------------------------------------------------------------------------------------
def int_to_le_bytes(v: int, nbytes: int) -> bytes:
    return int(v).to_bytes(nbytes, "little")


def le_bytes_to_int(b: bytes) -> int:
    return int.from_bytes(b, "little")


def fr_to_mont(a: int) -> int:
    return a * FR_R % R_MOD


def fr_from_mont(a: int) -> int:
    return a * FR_RINV % R_MOD


def fq_to_mont(a: int) -> int:
    return a * FQ_R % P_MOD


def fq_from_mont(a: int) -> int:
    return a * FQ_RINV % P_MOD


def fr_vec_to_bytes(vals, mont: bool = True) -> bytes:
    """n x 32 B, the in-memory image of Vec<Fp256<FrParameters>> (Montgomery) or of
    Vec<BigInteger256> (canonical, mont=False)."""
    out = bytearray()
    for v in vals:
        out += int_to_le_bytes(fr_to_mont(v) if mont else v, 32)
    return bytes(out)


def fr_vec_from_bytes(buf: bytes, mont: bool = True):
    n = len(buf) // 32
    vals = [le_bytes_to_int(buf[32 * i : 32 * i + 32]) for i in range(n)]
    return [fr_from_mont(v) for v in vals] if mont else vals


# ----------------------------------------------------------------------------------------------
# Deterministic inputs: SplitMix64 (language independent)
# ----------------------------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def next_fr(self) -> int:
        """uniform on [0, r): four u64 draws (limb 0 first), mask to 253 bits, reject >= r."""
        while True:
            v = 0
            for k in range(4):
                v |= self.next() << (64 * k)
            v &= (1 << 253) - 1
            if v < R_MOD:
                return v


def random_fr_vec(n: int, seed: int):
    g = SplitMix64(seed)
    return [g.next_fr() for _ in range(n)]


# ----------------------------------------------------------------------------------------------
# Fr radix-2 NTT (snarkVM EvaluationDomain semantics)
# ----------------------------------------------------------------------------------------------
def fr_root_of_unity(log_n: int) -> int:
    """omega = TWO_ADIC_ROOT_OF_UNITY squared (47 - log_n) times (Fr::get_root_of_unity)."""
    if log_n > FR_TWO_ADICITY:
        raise ValueError("domain too large")
    return pow(FR_TWO_ADIC_ROOT, 1 << (FR_TWO_ADICITY - log_n), R_MOD)


def _bitrev(i: int, bits: int) -> int:
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def _ntt_core(a, omega):
    """iterative Cooley-Tukey, natural in -> natural out, exact mod r."""
    n = len(a)
    log_n = n.bit_length() - 1
    a = list(a)
    for i in range(n):
        j = _bitrev(i, log_n)
        if i < j:
            a[i], a[j] = a[j], a[i]
    m = 1
    while m < n:
        w_m = pow(omega, n // (2 * m), R_MOD)
        tw = [1] * m
        for k in range(1, m):
            tw[k] = tw[k - 1] * w_m % R_MOD
        for s in range(0, n, 2 * m):
            for k in range(m):
                u = a[s + k]
                t = a[s + k + m] * tw[k] % R_MOD
                a[s + k] = (u + t) % R_MOD
                a[s + k + m] = (u - t) % R_MOD
        m *= 2
    return a


def domain_size(num_coeffs: int) -> int:
    """EvaluationDomain::new: next power of two; None (here: ValueError) above 2^47."""
    n = 1
    while n < num_coeffs:
        n *= 2
    if n.bit_length() - 1 > FR_TWO_ADICITY:
        raise ValueError("domain too large")
    return n


def fft(coeffs, n: int | None = None):
    """EvaluationDomain::fft: zero-pad to the domain size, evaluate at omega^k, natural order."""
    n = n or domain_size(len(coeffs))
    a = list(coeffs[:n]) + [0] * (n - min(len(coeffs), n))
    return _ntt_core(a, fr_root_of_unity(n.bit_length() - 1))


def ifft(evals, n: int | None = None):
    n = n or domain_size(len(evals))
    a = list(evals[:n]) + [0] * (n - min(len(evals), n))
    w_inv = pow(fr_root_of_unity(n.bit_length() - 1), -1, R_MOD)
    n_inv = pow(n, -1, R_MOD)
    return [v * n_inv % R_MOD for v in _ntt_core(a, w_inv)]


def coset_fft(coeffs, n: int | None = None):
    n = n or domain_size(len(coeffs))
    a = list(coeffs[:n]) + [0] * (n - min(len(coeffs), n))
    gp, out = 1, []
    for v in a:                       # distribute_powers(coeffs, g)
        out.append(v * gp % R_MOD)
        gp = gp * FR_GENERATOR % R_MOD
    return fft(out, n)


def coset_ifft(evals, n: int | None = None):
    a = ifft(evals, n)
    g_inv = pow(FR_GENERATOR, -1, R_MOD)
    gp, out = 1, []
    for v in a:
        out.append(v * gp % R_MOD)
        gp = gp * g_inv % R_MOD
    return out


def poly_eval(coeffs, x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc


def distribute_powers(coeffs, g: int, k: int = 1):
    """c_i <- c_i * k * g^i: EvaluationDomain::distribute_powers / distribute_powers_and_mul_by_const
    (snarkvm-algorithms 0.14.5 src/fft/domain.rs [U])."""
    out, p = [], k % R_MOD
    for c in coeffs:
        out.append(c * p % R_MOD)
        p = p * g % R_MOD
    return out


def lagrange_coefficients(log_n: int, tau: int):
    """EvaluationDomain::evaluate_all_lagrange_coefficients (snarkvm-algorithms 0.14.5 src/fft/domain.rs [U]):
    L_i(tau) = (tau^n - 1) / n * w^i / (tau - w^i); the indicator vector when tau is in the domain."""
    n = 1 << log_n
    w = fr_root_of_unity(log_n)
    tn = pow(tau, n, R_MOD)
    pw = [pow(w, i, R_MOD) for i in range(n)]
    if tn == 1:
        return [1 if p == tau % R_MOD else 0 for p in pw]
    c = (tn - 1) * pow(n, -1, R_MOD) % R_MOD
    return [c * p % R_MOD * pow((tau - p) % R_MOD, -1, R_MOD) % R_MOD for p in pw]


def divide_by_vanishing_on_coset(evals, log_n: int, g: int = None):
    """evals[i] = p(g w_m^i) over the coset of the domain of size m = len(evals), divided by Z_n(x) = x^n - 1,
    n = 2^log_n <= m (EvaluationDomain::divide_by_vanishing_poly_on_coset_in_place for m == n, snarkvm-algorithms
    0.14.5 src/fft/domain.rs [U]).  Every divisor is evaluated directly, no period shortcut."""
    g = FR_GENERATOR if g is None else g
    m = len(evals)
    w = fr_root_of_unity(m.bit_length() - 1)
    out, x = [], g % R_MOD
    for v in evals:
        d = (pow(x, 1 << log_n, R_MOD) - 1) % R_MOD
        out.append(v * pow(d, -1, R_MOD) % R_MOD if d else 0)
        x = x * w % R_MOD
    return out


def divide_by_linear(coeffs, z: int):
    """witness polynomial of KZG10::open: q(x) = (p(x) - p(z)) / (x - z) by synthetic division, returned with
    len(coeffs) entries (the last one 0) (KZG10::compute_witness_polynomial, src/polycommit/kzg10/mod.rs [U];
    SURVEY.md 8a row 13)."""
    n = len(coeffs)
    q = [0] * n
    carry = 0
    for i in range(n - 1, 0, -1):
        carry = (coeffs[i] + carry * z) % R_MOD
        q[i - 1] = carry
    return q


# ----------------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 1 over Fq.  Affine points are (x, y) tuples; None is the identity.
# ----------------------------------------------------------------------------------------------
def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - COEFF_B) % P_MOD == 0


def g1_neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % P_MOD)


def g1_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return (x3, (lam * (x1 - x3) - y1) % P_MOD)


# Jacobian helpers (X, Y, Z); identity Z == 0.  Used for speed inside scalar mul / MSM.
def _jac_double(P):
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return (0, 1, 0)
    A = X * X % P_MOD
    B = Y * Y % P_MOD
    C = B * B % P_MOD
    D = 2 * ((X + B) ** 2 - A - C) % P_MOD
    E = 3 * A % P_MOD
    F = E * E % P_MOD
    X3 = (F - 2 * D) % P_MOD
    Y3 = (E * (D - X3) - 8 * C) % P_MOD
    Z3 = 2 * Y * Z % P_MOD
    return (X3, Y3, Z3)


def _jac_add_affine(P, q):
    if q is None:
        return P
    X1, Y1, Z1 = P
    x2, y2 = q
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P_MOD
    U2 = x2 * Z1Z1 % P_MOD
    S2 = y2 * Z1 * Z1Z1 % P_MOD
    H = (U2 - X1) % P_MOD
    Rr = (S2 - Y1) % P_MOD
    if H == 0:
        if Rr == 0:
            return _jac_double(P)
        return (0, 1, 0)
    HH = H * H % P_MOD
    HHH = H * HH % P_MOD
    V = X1 * HH % P_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % P_MOD
    Y3 = (Rr * (V - X3) - Y1 * HHH) % P_MOD
    Z3 = Z1 * H % P_MOD
    return (X3, Y3, Z3)


def _jac_add(P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % P_MOD
    Z2Z2 = Z2 * Z2 % P_MOD
    U1 = X1 * Z2Z2 % P_MOD
    U2 = X2 * Z1Z1 % P_MOD
    S1 = Y1 * Z2 * Z2Z2 % P_MOD
    S2 = Y2 * Z1 * Z1Z1 % P_MOD
    H = (U2 - U1) % P_MOD
    Rr = (S2 - S1) % P_MOD
    if H == 0:
        if Rr == 0:
            return _jac_double(P)
        return (0, 1, 0)
    HH = H * H % P_MOD
    HHH = H * HH % P_MOD
    V = U1 * HH % P_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % P_MOD
    Y3 = (Rr * (V - X3) - S1 * HHH) % P_MOD
    Z3 = Z1 * Z2 * H % P_MOD
    return (X3, Y3, Z3)


def jac_to_affine(P):
    X, Y, Z = P
    if Z % P_MOD == 0:
        return None
    zi = pow(Z, -1, P_MOD)
    zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def g1_mul(pt, k: int):
    """k * pt, double-and-add on Jacobian coordinates (k taken mod nothing: any non-negative int)."""
    if pt is None or k == 0:
        return None
    acc = (0, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, pt)
    return jac_to_affine(acc)


def msm_naive(bases, scalars):
    """sum_i s_i * P_i by independent scalar multiplications (definition of VariableBase::msm;
    like upstream, the two slices are zipped: the shorter length wins)."""
    acc = (0, 1, 0)
    for pt, s in zip(bases, scalars):
        q = g1_mul(pt, s)
        acc = _jac_add_affine(acc, q)
    return jac_to_affine(acc)


def msm_pippenger(bases, scalars, c: int | None = None):
    """Independent second algorithm (unsigned-window bucket method) used to cross-check msm_naive
    and to serve as the oracle at sizes where msm_naive is too slow."""
    n = min(len(bases), len(scalars))
    if n == 0:
        return None
    if c is None:
        c = max(2, min(16, n.bit_length() - 2))
    windows = (253 + c - 1) // c
    total = (0, 1, 0)
    for w in reversed(range(windows)):
        for _ in range(c):
            total = _jac_double(total)
        buckets = [(0, 1, 0)] * (1 << c)
        for i in range(n):
            d = (scalars[i] >> (w * c)) & ((1 << c) - 1)
            if d and bases[i] is not None:
                buckets[d] = _jac_add_affine(buckets[d], bases[i])
        run = (0, 1, 0)
        acc = (0, 1, 0)
        for d in range((1 << c) - 1, 0, -1):
            run = _jac_add(run, buckets[d])
            acc = _jac_add(acc, run)
        total = _jac_add(total, acc)
    return jac_to_affine(total)


# ----------------------------------------------------------------------------------------------
# Synthetic bases with known discrete logs:  P_i = (s0 + i*d) * G   
# so that  msm(P, s) == (sum_i s_i * (s0 + i*d) mod r) * G  can be checked at ANY size.
# ----------------------------------------------------------------------------------------------
def base_dlogs(n: int, seed: int):
    g = SplitMix64(seed ^ 0xB200B200)
    s0, d = g.next_fr(), g.next_fr()
    return s0, d


def synthetic_bases(n: int, seed: int):
    s0, d = base_dlogs(n, seed)
    p0 = g1_mul(G1_GEN, s0)
    q = g1_mul(G1_GEN, d)
    out, cur = [], p0
    for _ in range(n):
        out.append(cur)
        cur = g1_add(cur, q)
    return out


def msm_expected_from_dlogs(n: int, seed: int, scalars):
    s0, d = base_dlogs(n, seed)
    k = 0
    for i, s in enumerate(scalars[:n]):
        k += s * (s0 + i * d)
    return g1_mul(G1_GEN, k % R_MOD)


# ----------------------------------------------------------------------------------------------
# Memory images crossing the C ABI
# ----------------------------------------------------------------------------------------------
AFFINE_STRIDE_RUST = 104   # size_of::<G1Affine>() : x(48) y(48) infinity(1) pad(7)
AFFINE_STRIDE_PACKED = 96


def g1_affine_to_bytes(pt, stride: int = AFFINE_STRIDE_RUST) -> bytes:
    if pt is None:   # snarkVM: Affine::zero() = (0, 1, infinity = true)
        x, y, inf = 0, 1, 1
    else:
        x, y, inf = pt[0], pt[1], 0
    b = int_to_le_bytes(fq_to_mont(x), 48) + int_to_le_bytes(fq_to_mont(y), 48)
    if stride == AFFINE_STRIDE_PACKED:
        if inf:     # packed form has no flag: the identity is encoded as (0, 0), which is off-curve
            b = bytes(96)
        return b
    return b + bytes([inf]) + bytes(stride - 97)


def g1_affine_vec_to_bytes(pts, stride: int = AFFINE_STRIDE_RUST) -> bytes:
    return b"".join(g1_affine_to_bytes(p, stride) for p in pts)


def g1_affine_from_bytes(buf: bytes, stride: int = AFFINE_STRIDE_RUST):
    x = fq_from_mont(le_bytes_to_int(buf[0:48]))
    y = fq_from_mont(le_bytes_to_int(buf[48:96]))
    if stride == AFFINE_STRIDE_RUST:
        if buf[96]:
            return None
    elif x == 0 and y == 0:
        return None
    return (x, y)


def g1_projective_from_bytes(buf: bytes):
    """144-byte Jacobian {x, y, z} in Montgomery form -> affine tuple / None."""
    X = fq_from_mont(le_bytes_to_int(buf[0:48]))
    Y = fq_from_mont(le_bytes_to_int(buf[48:96]))
    Z = fq_from_mont(le_bytes_to_int(buf[96:144]))
    return jac_to_affine((X, Y, Z))


def g1_projective_to_bytes(pt) -> bytes:
    """affine tuple / None -> normalised Jacobian image (Z = 1, or (0, 1, 0) for the identity)."""
    if pt is None:
        X, Y, Z = 0, 1, 0
    else:
        X, Y, Z = pt[0], pt[1], 1
    return b"".join(int_to_le_bytes(fq_to_mont(v), 48) for v in (X, Y, Z))


# ----------------------------------------------------------------------------------------------
# Wire formats pinned by the reference's proof fixture
# ----------------------------------------------------------------------------------------------
def fq_sqrt(a: int):
    """Tonelli-Shanks over Fq (two-adicity 46, QNR generator 15)."""
    a %= P_MOD
    if a == 0:
        return 0
    if pow(a, (P_MOD - 1) // 2, P_MOD) != 1:
        return None
    s, q = 0, P_MOD - 1
    while q % 2 == 0:
        q //= 2
        s += 1
    z = pow(15, q, P_MOD)
    m, c, t, rr = s, z, pow(a, q, P_MOD), pow(a, (q + 1) // 2, P_MOD)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % P_MOD
            i += 1
        b = pow(c, 1 << (m - i - 1), P_MOD)
        m, c = i, b * b % P_MOD
        t, rr = t * c % P_MOD, rr * b % P_MOD
    return rr


def g1_decompress(buf: bytes):
    """48-byte compressed G1: x little-endian with two flag bits in the top byte
    (bit 7: y is the lexicographically larger root; bit 6: infinity)."""
    assert len(buf) == 48
    v = le_bytes_to_int(buf)
    y_flag = (v >> 383) & 1
    inf_flag = (v >> 382) & 1
    x = v & ((1 << 382) - 1)
    if inf_flag:
        return None
    if x >= P_MOD:
        raise ValueError("x out of range")
    y = fq_sqrt((x * x * x + COEFF_B) % P_MOD)
    if y is None:
        raise ValueError("x not on curve")
    larger = y > (P_MOD - 1) // 2
    if larger != bool(y_flag):
        y = P_MOD - y
    return (x, y)


def curve_points_outside_subgroup(count: int):
    """points on y^2 = x^3 + 1 that are NOT in the r-torsion (small x; r * P != O checked) -- negative fixtures for
    deserialize_compressed's subgroup check"""
    out, x = [], 2
    while len(out) < count:
        y = fq_sqrt((x * x * x + COEFF_B) % P_MOD)
        if y is not None and g1_mul((x, y), R_MOD) is not None:
            out.append((x, y))
        x += 1
    return out


def g1_compress(pt) -> bytes:
    if pt is None:
        return int_to_le_bytes(1 << 382, 48)
    x, y = pt
    flag = 1 if y > (P_MOD - 1) // 2 else 0
    return int_to_le_bytes(x | (flag << 383), 48)


_BECH32_CHARSET = "qpzry9x8gf2tvdw0s3jn54khce6mua7l"
_BECH32M_CONST = 0x2BC830A3


def _bech32_polymod(values):
    gen = [0x3B6A57B2, 0x26508E6D, 0x1EA119FA, 0x3D4233DD, 0x2A1462B3]
    chk = 1
    for v in values:
        b = chk >> 25
        chk = ((chk & 0x1FFFFFF) << 5) ^ v
        for i in range(5):
            chk ^= gen[i] if ((b >> i) & 1) else 0
    return chk


def bech32m_decode(s: str):
    """-> (hrp, payload bytes).  Raises on checksum failure."""
    pos = s.rfind("1")
    hrp, data = s[:pos], [_BECH32_CHARSET.index(ch) for ch in s[pos + 1 :]]
    exp = [ord(ch) >> 5 for ch in hrp] + [0] + [ord(ch) & 31 for ch in hrp]
    if _bech32_polymod(exp + data) != _BECH32M_CONST:
        raise ValueError("bad bech32m checksum")
    acc, bits, out = 0, 0, bytearray()
    for v in data[:-6]:
        acc = (acc << 5) | v
        bits += 5
        while bits >= 8:
            bits -= 8
            out.append((acc >> bits) & 0xFF)
    return hrp, bytes(out)
