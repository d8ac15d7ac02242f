"""The mathematical rearrangements the fast kernels rest on (DESIGN.md sections 2-4), restated in numpy / Python
integers and checked against the oracle on CPU.  These are models of the kernels' arithmetic, not the kernels: they pin
the claims "this parallel form is bit-identical to the reference's sequential one" independently of any GPU run, and
they fail here -- where the reference tree and the oracle are at hand -- if a form is ever changed.

  * fir_q15 / polydec_q15 / upfir: byte-plane DP2A accumulation is exact in int32 for 256 taps per plane
  * cic_dec_fast (R=8, N=4, M=1): three decimating [1 4 6 4 1] half-band stages with phases (1, 0, -1)
  * cic_intr_fast: polyphase boxcar(RM)^N form, the N-1 discard and the stream-edge output count
  * upfir (cicfir): composite taps boxcar(RM)^N * g make the two reference objects one integer FIR modulo 2^W
  * CIC history carry: integrators restarted from zero N*M low-rate samples early are annihilated by the N combs
"""
import numpy as np
import pytest

from oracle import oracle as O

Q15, ACC40 = (16, 1), (40, 8)


def wrap(v, W):
    """Python-int / object-array two's-complement wrap to W bits."""
    m = (1 << W) - 1
    v = np.asarray(v, dtype=object) & m
    return np.where(v >> (W - 1) & 1, v - (1 << W), v)


def boxcar_pow(L, N):
    h = np.array([1], dtype=object)
    for _ in range(N):
        h = np.convolve(h, np.ones(L, dtype=object))
    return h


def rand16(rng, n):
    return rng.integers(-32768, 32767, size=n, endpoint=True).astype(np.int16)


# ---------------------------------------------------------------------------------------------- byte planes
@pytest.mark.parametrize("data", ["random", "all_min_x_ff", "all_min_x_min", "alternating"])
def test_byte_plane_accumulation_is_exact_in_int32_for_256_taps(data):
    """h = 256*hh + hl (hl unsigned byte, hh signed byte): sum x*hl and sum x*hh over 256 taps each fit an int32, so
    dp2a accumulation per plane is exact and (hi << 8) + lo is the 16x16 dot product (fir_q15.cu header)."""
    rng = np.random.default_rng(5)
    n = 256
    if data == "random":
        x, h = rand16(rng, n), rand16(rng, n)
    elif data == "all_min_x_ff":                    # worst case of the low plane: |x| = 2^15, hl = 255
        x, h = np.full(n, -32768, np.int16), np.full(n, 0x00FF, np.int16)
    elif data == "all_min_x_min":                   # worst case of the high plane: hh = -128
        x, h = np.full(n, -32768, np.int16), np.full(n, -32768, np.int16)
    else:
        x = np.where(np.arange(n) & 1, 32767, -32768).astype(np.int16)
        h = np.where(np.arange(n) & 1, -1, 0x7FFF).astype(np.int16)
    hl = (h.astype(np.int32) & 0xFF).astype(np.int32)
    hh = (h.astype(np.int32) >> 8).astype(np.int32)
    lo32 = np.int32(0)
    hi32 = np.int32(0)
    with np.errstate(over="ignore"):
        for k in range(0, n, 2):                    # one dp2a per plane and tap pair, int32 wrap-around adds
            lo32 = np.int32(lo32 + np.int32(x[k]) * hl[k] + np.int32(x[k + 1]) * hl[k + 1])
            hi32 = np.int32(hi32 + np.int32(x[k]) * hh[k] + np.int32(x[k + 1]) * hh[k + 1])
    exact = int(np.dot(x.astype(np.int64), h.astype(np.int64)))
    assert (int(hi32) << 8) + int(lo32) == exact
    assert abs(int(np.dot(x.astype(np.int64), hl.astype(np.int64)))) < 2 ** 31
    assert abs(int(np.dot(x.astype(np.int64), hh.astype(np.int64)))) <= 2 ** 31 - 1 or data == "all_min_x_min"


def test_three_byte_planes_cover_24_bit_composite_taps():
    """upfir_q15_pack: c = b0 + 256*b1 + 65536*b2 with b0, b1 unsigned bytes and b2 a signed byte; per plane <= 128 taps
    per phase stay inside an int32."""
    rng = np.random.default_rng(6)
    c = rng.integers(-(1 << 23), (1 << 23) - 1, size=128, endpoint=True)
    x = rand16(rng, 128).astype(np.int64)
    b0, b1, b2 = c & 0xFF, (c >> 8) & 0xFF, c >> 16
    assert np.all((b2 >= -128) & (b2 <= 127))
    p = [int(np.dot(x, b)) for b in (b0, b1, b2)]
    assert all(abs(v) < 2 ** 31 for v in p)
    assert p[0] + (p[1] << 8) + (p[2] << 16) == int(np.dot(x, c))
    worst = 128 * 32768 * 255
    assert worst < 2 ** 31


# ------------------------------------------------------------------------------------- CIC decimator, non-recursive
def halfband_decimate(v, phi):
    """w[q] = sum_t b[t] * v[2q + phi - t], b = [1 4 6 4 1], v[<0] = 0; as many q as the input supports causally."""
    b = [1, 4, 6, 4, 1]
    nq = (len(v) - phi + 1) // 2 if len(v) else 0           # 2q + phi <= len(v) - 1
    w = np.zeros(max(nq, 0), dtype=object)
    for q in range(len(w)):
        s = 0
        for t, bt in enumerate(b):
            i = 2 * q + phi - t
            if 0 <= i < len(v):
                s += bt * int(v[i])
        w[q] = s
    return w


@pytest.mark.parametrize("n", [1, 7, 8, 9, 64, 1000, 4099])
def test_cic_dec_r8_n4_equals_three_halfband_stages(n):
    """cic_fast.cu: boxcar(8)^4 = prod_s (1 + z^-2^s)^4; with phases (1, 0, -1) the third decimating stage is the
    reference's out[m] = (boxcar(8)^4 * x)[8m - 3] (ac_cic_full_core.h:80-135,198-255)."""
    rng = np.random.default_rng(n)
    x = rand16(rng, n)
    ref = O.CicB("dec", Q15, (28, 13), 8, 1, 4).run(x)
    # closed form first, then the factorised evaluation
    h = boxcar_pow(8, 4)
    full = np.convolve(x.astype(object), h)
    closed = np.array([full[8 * m - 3] if 8 * m - 3 >= 0 else 0 for m in range(len(ref))], dtype=object)
    assert np.array_equal(wrap(closed, 28).astype(np.int64), ref)
    xp = np.concatenate([x.astype(object), np.zeros(32, dtype=object)])      # zeros past the end: only causal outputs are compared
    v = xp
    for phi in (1, 0, -1):
        v = halfband_decimate(v, phi)
    assert len(v) >= len(ref)
    assert np.array_equal(wrap(v[:len(ref)], 28).astype(np.int64), ref)


# ------------------------------------------------------------------------------------ CIC interpolator, polyphase
@pytest.mark.parametrize("R,M,N", [(4, 1, 3), (2, 1, 4), (8, 2, 3), (7, 2, 5)])
@pytest.mark.parametrize("K", [1, 2, 5, 40, 333])
def test_cic_intr_equals_polyphase_boxcar_form(R, M, N, K):
    """cic_intr_fast.cu: out[kR + ph] = sum_m h[ph + R m] x[k - m], h = boxcar(RM)^N; the pipelined integrator's N-1
    samples of latency and the N-1 discarded outputs cancel; K inputs give max(0, (K-1)R + 1 - (N-1)) outputs
    (ac_cic_intr_full.h:150-215, ac_cic_full_core.h:143-160)."""
    rng = np.random.default_rng(K * 31 + R)
    W = O.cic_int_width("intr", Q15, R, M, N)
    outf = (W, W - 15)                                             # the lossless internal type, passed on unchanged
    x = rand16(rng, K)
    ref = O.CicB("intr", Q15, outf, R, M, N).run(x)
    assert len(ref) == max(0, (K - 1) * R + 1 - (N - 1))
    h = boxcar_pow(R * M, N)
    out = np.zeros(len(ref), dtype=object)
    for j in range(len(ref)):
        k, ph = divmod(j, R)
        s = 0
        m = 0
        while ph + R * m < len(h) and k - m >= 0:
            s += int(h[ph + R * m]) * int(x[k - m])
            m += 1
        out[j] = s
    assert np.array_equal(wrap(out, W).astype(np.int64), ref)


# ---------------------------------------------------------------------------------------- fused cascade (cicfir)
@pytest.mark.parametrize("taps", [1, 8, 63])
def test_cascade_equals_one_fir_with_composite_taps(taps):
    """upfir_q15.cu: the lossless CIC type passed on unchanged and an exact-shift AC_WRAP accumulator make
    ac_cic_intr_full -> ac_fir_* one integer FIR with taps boxcar(RM)^N * g, modulo 2^W_acc (BASELINE configs[4])."""
    R, M, N, mid = 4, 1, 3, (20, 5)
    rng = np.random.default_rng(taps)
    x = rand16(rng, 400)
    g = rand16(rng, taps)
    cic = O.CicB("intr", Q15, mid, R, M, N)
    fir = O.FirB(mid, Q15, ACC40, ACC40, taps, "SHIFT_REG")
    fir.load(g)
    ref = fir.run(cic.run(x))
    c = np.convolve(boxcar_pow(R * M, N), g.astype(object))
    assert max(abs(int(v)) for v in c) < 1 << 23                     # three byte planes
    lsh = (40 - 8) - ((20 - 5) + 15)                                 # F_acc - (F_mid + F_coeff) = 2: exact left shift
    out = np.zeros(len(ref), dtype=object)
    for j in range(len(ref)):
        k, ph = divmod(j, R)
        s, m = 0, 0
        while ph + R * m < len(c) and k - m >= 0:
            s += int(c[ph + R * m]) * int(x[k - m])
            m += 1
        out[j] = s << lsh
    assert np.array_equal(wrap(out, 40).astype(np.int64), ref)


# -------------------------------------------------------------------------------- history carry of the CIC decimator
@pytest.mark.parametrize("R,M,N", [(8, 1, 4), (7, 2, 4), (2, 1, 5), (16, 1, 3)])
def test_cic_dec_restart_from_zero_is_annihilated_by_the_combs(R, M, N):
    """the only state carried between calls is the last H = N*M*R + N - 1 inputs.  Integrators
    restarted from zero that far back differ from the true ones by a polynomial of degree N-1 in the low-rate index,
    which N combs of delay M remove exactly: a fresh object fed x[s-H':] (H' = H rounded up to a multiple of R so the
    decimation phase is kept) reproduces the tail of the full run after N*M + 1 warm-up outputs."""
    rng = np.random.default_rng(R * 100 + N)
    x = rand16(rng, 6000)
    W = O.cic_int_width("dec", Q15, R, M, N)
    outf = (W, W - 15)
    full = O.CicB("dec", Q15, outf, R, M, N).run(x)
    H = N * M * R + N - 1
    Hp = -(-H // R) * R
    for s in (R * 40, R * 333, R * 700):
        part = O.CicB("dec", Q15, outf, R, M, N).run(x[s - Hp:])
        skip = Hp // R
        assert skip >= N * M + 1
        assert np.array_equal(part[skip:], full[s // R:]), (R, M, N, s)


# ----------------------------------------------------------------------------------------- polyphase decimator
@pytest.mark.parametrize("NT,DF", [(32, 8), (4, 2), (7, 3), (1, 4)])
def test_poly_dec_equals_sum_of_phase_firs(NT, DF):
    """fir_dec.cu: out[m] = sum_r sum_tp coeffs[tp + NTAPS*r] * u_r[m - tp], u_r[m] = x[m*DF + DF-1 - r]
    (ac_poly_dec.h:101-131: phases DF-1 .. 0, one output per DF inputs); <16,1> x <16,1> -> <40,8> is an exact left
    shift by 2, so the sum is order-free modulo 2^40."""
    rng = np.random.default_rng(NT * 10 + DF)
    x = rand16(rng, DF * 200 + DF - 1)               # a ragged tail: the last DF-1 inputs produce nothing yet
    h = rand16(rng, NT * DF)
    o = O.PdB(Q15, Q15, ACC40, ACC40, NT, DF)
    o.load(h)
    ref = o.run(x)
    assert len(ref) == len(x) // DF
    out = np.zeros(len(ref), dtype=object)
    for m in range(len(ref)):
        s = 0
        for r in range(DF):
            for tp in range(NT):
                i = (m - tp) * DF + DF - 1 - r
                if i >= 0:
                    s += int(h[tp + NT * r]) * int(x[i])
        out[m] = s << 2
    assert np.array_equal(wrap(out, 40).astype(np.int64), ref)


# ------------------------------------------------------------------------- polyphase interpolator, plain form
@pytest.mark.parametrize("NT,IF", [(16, 4), (3, 2), (5, 8)])
def test_poly_intr_plain_form_equals_upsampling_fir(NT, IF):
    """upfir_lane with two byte planes serves ac_poly_intr FOLD_ANTI (ac_poly_intr.h:232-251):
    out[k*IF + j] = sum_i coeffs[i + NTAPS*j] * x[k - i], written at once."""
    rng = np.random.default_rng(NT + IF)
    x = rand16(rng, 300)
    h = rand16(rng, NT * IF)
    o = O.PiB(Q15, Q15, ACC40, ACC40, NT, IF, "FOLD_ANTI")
    o.load(h)
    ref = o.run(x)
    assert len(ref) == len(x) * IF
    out = np.zeros(len(ref), dtype=object)
    for n in range(len(ref)):
        k, j = divmod(n, IF)
        out[n] = sum(int(h[i + NT * j]) * int(x[k - i]) for i in range(NT) if k - i >= 0) << 2
    assert np.array_equal(wrap(out, 40).astype(np.int64), ref)


# ------------------------------------------------------------------------------------------ integrate-and-dump
@pytest.mark.parametrize("CHN", [1, 2, 4, 8])
def test_intg_dump_equal_frames_are_segment_sums(CHN):
    """intgdump_vec: with every n_sample token equal to ns, dump d of channel c is the sum of x[(d*ns + i)*CHN + c]
    (ac_intg_dump.h:96-149); <16,1> -> <32,17> keeps every bit (F_acc = F_in)."""
    rng = np.random.default_rng(CHN)
    ns, frames = 64, 12
    x = rand16(rng, ns * frames * CHN)
    ref = O.IdB(Q15, (32, 17), (32, 17), 1024, CHN).run(x, np.full(frames, ns))
    seg = x.astype(np.int64).reshape(frames, ns, CHN).sum(axis=1)
    assert np.array_equal(np.asarray(ref).reshape(frames, CHN), seg)


# ----------------------------------------------------------------- the reference's comb for M > 2 (a quirk, followed)
@pytest.mark.parametrize("mode,R,M,N", [("dec", 4, 3, 2), ("dec", 2, 5, 3), ("intr", 3, 3, 2), ("intr", 5, 4, 3), ("dec", 8, 3, 1)])
def test_comb_delay_of_the_reference_is_min_m_2(mode, R, M, N):
    """ac_cic_full_core.h:247-251 shifts the comb delay line with an ascending copy loop, so its differential delay is
    min(M, 2) while the lossless width still grows with M (oracle_b.c cic_comb, runtime.cu cic_comb_delay).  The engine
    therefore runs an M > 2 instantiation as the M = 2 filter at the M-wide internal type: same outputs."""
    rng = np.random.default_rng(M * 7 + N)
    x = rand16(rng, 500)
    W = O.cic_int_width(mode, Q15, R, M, N)
    outf = (W, W - 15)
    y_m = O.CicB(mode, Q15, outf, R, M, N).run(x)
    y_2 = O.CicB(mode, Q15, outf, R, 2, N).run(x)
    assert np.array_equal(y_m, y_2)
    assert not np.array_equal(y_m, O.CicB(mode, Q15, outf, R, 1, N).run(x)) or N == 0
