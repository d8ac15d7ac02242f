/* oracle/oracle_b.c -- TEST INFRASTRUCTURE (Oracle B), not product code.
 *
 * Header-free CPU restatement, on raw two's-complement integers, of the reference's FIR
 * and CIC run() paths.  Every function cites the reference lines it restates
 * (paths relative to /root/reference/include/ac_dsp/).  The ac_fixed arithmetic itself
 * lives in hlslibs/ac_types (not vendored by the reference, unpinned -- SURVEY.md 8c);
 * its published rules are restated in ob_convert()/ob_macc():
 *   product : exact, F = F1 + F2
 *   sum     : exact, F = max(F1, F2)
 *   assign  : drop fraction bits with quantisation mode Q, then integer bits with overflow mode O
 *   a += b  : a = convert(a + b)
 * Pinned against: the five reference benches via Oracle A (reference headers over the
 * clean-room shim), the two CIC golden vectors, and randomised A == B sweeps
 * (tests/test_oracle_*.py).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 w128;
typedef unsigned __int128 u128;

typedef struct { int W, I, S, Q, O; } ob_fmt;  /* ac_fixed<W,I,S,Q,O>; Q/O use the ac_q_mode / ac_o_mode order */
enum { Q_TRN, Q_RND, Q_TRN_ZERO, Q_RND_ZERO, Q_RND_INF, Q_RND_MIN_INF, Q_RND_CONV, Q_RND_CONV_ODD };
enum { O_WRAP, O_SAT, O_SAT_ZERO, O_SAT_SYM };
enum { FT_SHIFT_REG, FT_ROTATE_SHIFT, FT_C_BUFF, FT_FOLD_EVEN, FT_FOLD_ODD, FT_TRANSPOSED, FT_FOLD_EVEN_ANTI, FT_FOLD_ODD_ANTI };

static int F_of(const ob_fmt *f) { return f->W - f->I; }

static w128 ob_wrap(w128 v, int W, int S) {
  if (W >= 128) return v;
  u128 m = (((u128)1) << W) - 1;
  u128 u = ((u128)v) & m;
  if (S && ((u >> (W - 1)) & 1)) u |= ~m;
  return (w128)u;
}

/* value with F2 fraction bits -> format f (quantise, then overflow) */
static w128 ob_convert(w128 v, int F2, const ob_fmt *f) {
  int F = F_of(f);
  if (F2 > F) {
    int sh = F2 - F;
    w128 q = v >> sh;
    w128 rem = v - (q << sh);
    int msb = (int)((rem >> (sh - 1)) & 1);
    int rest = (rem & ((((w128)1) << (sh - 1)) - 1)) != 0;
    int neg = v < 0;
    switch (f->Q) {
      case Q_TRN: break;
      case Q_RND: q += msb; break;
      case Q_TRN_ZERO: q += (neg && rem != 0); break;
      case Q_RND_INF: q += (msb && (rest || !neg)); break;
      case Q_RND_ZERO: q += (msb && (rest || neg)); break;
      case Q_RND_MIN_INF: q += (msb && rest); break;
      case Q_RND_CONV: q += (msb && (rest || (q & 1))); break;
      case Q_RND_CONV_ODD: q += (msb && (rest || !(q & 1))); break;
    }
    v = q;
  } else if (F > F2) {
    v = v << (F - F2);
  }
  {
    w128 hi = f->S ? ((((w128)1) << (f->W - 1)) - 1) : ((((w128)1) << f->W) - 1);
    w128 lo = f->S ? -(((w128)1) << (f->W - 1)) : 0;
    switch (f->O) {
      case O_WRAP: return ob_wrap(v, f->W, f->S);
      case O_SAT: return v > hi ? hi : (v < lo ? lo : v);
      case O_SAT_ZERO: return (v > hi || v < lo) ? 0 : v;
      case O_SAT_SYM: { w128 slo = f->S ? -hi : 0; return v > hi ? hi : (v < slo ? slo : v); }
    }
  }
  return v;
}

/* acc (format fa) += p, where p has Fp fraction bits: exact sum at max(F) then assign. */
static w128 ob_macc(w128 acc, const ob_fmt *fa, w128 p, int Fp) {
  int Fa = F_of(fa);
  int rF = Fa > Fp ? Fa : Fp;
  w128 s = (acc << (rF - Fa)) + (p << (rF - Fp));
  return ob_convert(s, rF, fa);
}

/* ------------------------------------------------------------------ FIR */
typedef struct {
  ob_fmt in, coeff, acc, out;
  int n, ftype;
  w128 *reg;        /* IN_TYPE reg[N_TAPS]         ac_fir_load_coeffs.h:127 */
  w128 *reg_trans;  /* ACC_TYPE reg_trans[N_TAPS]  ac_fir_load_coeffs.h:128 */
  w128 *h;          /* COEFF_TYPE coeffs[N_TAPS]   ac_fir_load_coeffs.h:306 */
  long wptr;        /* ac_fir_load_coeffs.h:129 */
} ob_fir;

ob_fir *ob_fir_create(const ob_fmt *in, const ob_fmt *coeff, const ob_fmt *acc, const ob_fmt *out, int n_taps, int ftype) {
  ob_fir *f = (ob_fir *)calloc(1, sizeof(ob_fir));
  f->in = *in; f->coeff = *coeff; f->acc = *acc; f->out = *out;
  f->n = n_taps; f->ftype = ftype;
  /* constructors zero the delay lines: ac_fir_load_coeffs.h:134-139 */
  f->reg = (w128 *)calloc(n_taps, sizeof(w128));
  f->reg_trans = (w128 *)calloc(n_taps, sizeof(w128));
  f->h = (w128 *)calloc(n_taps, sizeof(w128));
  f->wptr = 0;
  return f;
}
void ob_fir_destroy(ob_fir *f) { if (f) { free(f->reg); free(f->reg_trans); free(f->h); free(f); } }

/* load phase of ac_fir_load_coeffs::run (ac_fir_load_coeffs.h:324-331), the ctor pointer of
 * ac_fir_const_coeffs (ac_fir_const_coeffs.h:314) and the array argument of
 * ac_fir_prog_coeffs::run (ac_fir_prog_coeffs.h:277): N_TAPS raw coefficients. */
void ob_fir_load(ob_fir *f, const int64_t *c) {
  for (int i = 0; i < f->n; i++) f->h[i] = ob_wrap((w128)c[i], f->coeff.W, f->coeff.S);
}

/* firShiftReg: ac_fir_load_coeffs.h:145-151 */
static void fir_shift(ob_fir *f, w128 din) {
  for (int i = f->n - 1; i >= 0; i--) f->reg[i] = (i == 0) ? din : f->reg[i - 1];
}
/* firCircularBuffWrite / Read: ac_fir_load_coeffs.h:157-174 */
static void fir_cb_write(ob_fir *f, w128 din) {
  f->reg[f->wptr] = din;
  if (f->wptr == f->n - 1) f->wptr = 0; else f->wptr++;
}
static w128 fir_cb_read(ob_fir *f, long idx) {
  long rptr = f->wptr - 1 - idx;
  if (rptr < 0) rptr += f->n;
  return f->reg[rptr];
}

static w128 fir_one(ob_fir *f, w128 x) {
  const int N = f->n;
  const int Fin = F_of(&f->in), Fc = F_of(&f->coeff), Fa = F_of(&f->acc);
  w128 acc = 0;
  switch (f->ftype) {
    case FT_SHIFT_REG:  /* fir*ShiftReg: ac_fir_load_coeffs.h:180-188 */
      fir_shift(f, x);
      for (int i = N - 1; i >= 0; i--) acc = ob_macc(acc, &f->acc, f->reg[i] * f->h[i], Fin + Fc);
      break;
    case FT_ROTATE_SHIFT: {  /* fir*RotateShift: ac_fir_load_coeffs.h:194-208 */
      w128 t;
      for (int i = N; i >= 0; i--) {
        if (i == N) t = x;
        else { t = f->reg[N - 1]; acc = ob_macc(acc, &f->acc, f->reg[N - 1] * f->h[i], Fin + Fc); }
        fir_shift(f, t);
      }
      break;
    }
    case FT_C_BUFF:  /* fir*CircularBuff: ac_fir_load_coeffs.h:214-224 */
      for (int i = 0; i <= N - 1; i++) {
        if (i == 0) fir_cb_write(f, x);
        acc = ob_macc(acc, &f->acc, fir_cb_read(f, i) * f->h[i], Fin + Fc);
      }
      break;
    case FT_FOLD_EVEN:  /* fir*SymmetricEvenTaps: ac_fir_load_coeffs.h:231-239 (pre-add exact) */
      fir_shift(f, x);
      for (int i = N / 2 - 1; i >= 0; i--) acc = ob_macc(acc, &f->acc, f->h[i] * (f->reg[i] + f->reg[N - 1 - i]), Fin + Fc);
      break;
    case FT_FOLD_ODD:  /* fir*SymmetricOddTaps: ac_fir_load_coeffs.h:246-259 (`fold` is ACC_TYPE) */
      fir_shift(f, x);
      for (int i = 0; i < (N - 1) / 2 + 1; i++) {
        w128 fold;
        if (i == (N - 1) / 2) fold = ob_convert(f->reg[i], Fin, &f->acc);
        else fold = ob_convert(f->reg[i] + f->reg[N - 1 - i], Fin, &f->acc);
        acc = ob_macc(acc, &f->acc, f->h[i] * fold, Fc + Fa);
      }
      break;
    case FT_TRANSPOSED: {  /* fir*Transposed: ac_fir_load_coeffs.h:265-278 */
      for (int i = N - 1; i >= 0; i--) {
        w128 temp = (i == 0) ? 0 : f->reg_trans[i - 1];
        f->reg_trans[i] = ob_macc(temp, &f->acc, x * f->h[N - 1 - i], Fin + Fc);
      }
      acc = f->reg_trans[N - 1];
      break;
    }
    default: return 0;
  }
  return ob_convert(acc, Fa, &f->out);  /* data_out = acc: ac_fir_load_coeffs.h:187 */
}

/* sample loop of run(): ac_fir_load_coeffs.h:335-364 / ac_fir_const_coeffs.h:325-354 /
 * ac_fir_prog_coeffs.h:281-302 called once per sample. n inputs -> n outputs. */
long ob_fir_run(ob_fir *f, const int64_t *in, long n, int64_t *out) {
  if (f->ftype < 0 || f->ftype > FT_TRANSPOSED) return -1;
  for (long k = 0; k < n; k++) out[k] = (int64_t)fir_one(f, ob_wrap((w128)in[k], f->in.W, f->in.S));
  return n;
}

/* ------------------------------------------------------------------ ac_fir_reg_share */
/* One object of include/ac_dsp/ac_fir_reg_share.h:257-303 with its delay line (reg[N_TAPS], zero-initialised here;
 * caller-owned in the reference).  run() takes ONE sample and the coefficient RAM per call. */
typedef struct {
  ob_fmt in, coeff, acc, out;
  int n, ftype, mww, blk_sz, blk_off;
  w128 *reg;
} ob_rs;

ob_rs *ob_rs_create(const ob_fmt *in, const ob_fmt *coeff, const ob_fmt *acc, const ob_fmt *out, int n_taps, int mww, int blk_sz,
                    int blk_off, int ftype) {
  ob_rs *f = (ob_rs *)calloc(1, sizeof(ob_rs));
  f->in = *in; f->coeff = *coeff; f->acc = *acc; f->out = *out;
  f->n = n_taps; f->ftype = ftype; f->mww = mww; f->blk_sz = blk_sz; f->blk_off = blk_off;
  f->reg = (w128 *)calloc(n_taps, sizeof(w128));
  return f;
}
void ob_rs_destroy(ob_rs *f) { if (f) { free(f->reg); free(f); } }

/* words of coefficient RAM the block loops touch */
int ob_rs_ram_words(const ob_rs *f) {
  int used = f->ftype == FT_SHIFT_REG ? f->n : ((f->ftype == FT_FOLD_EVEN || f->ftype == FT_FOLD_EVEN_ANTI) ? f->n / 2 : (f->n - 1) / 2 + 1);
  int nblk = (used + f->blk_sz - 1) / f->blk_sz;
  return (nblk - 1) * f->mww + f->blk_off + f->blk_sz;
}

static w128 rs_one(ob_rs *f, w128 x, const w128 *ram) {
  const int N = f->n;
  const int Fin = F_of(&f->in), Fc = F_of(&f->coeff), Fa = F_of(&f->acc);
  w128 acc = 0;
  int ram_addr = 0;
  /* firShiftReg: ac_fir_reg_share.h:105-111 */
  for (int i = N - 1; i >= 0; i--) f->reg[i] = (i == 0) ? x : f->reg[i - 1];
  switch (f->ftype) {
    case FT_SHIFT_REG:       /* firProgCoeffsShiftReg: :122-135, taps visited UPWARDS, blocked RAM addressing */
      for (int i = 0; i < N; i += f->blk_sz, ram_addr += f->mww)
        for (int bc = f->blk_off, index = 0; bc < f->blk_off + f->blk_sz; bc++, index++)
          acc = ob_macc(acc, &f->acc, f->reg[i + index] * ram[ram_addr + bc], Fin + Fc);
      break;
    case FT_FOLD_EVEN:       /* ...SymmetricEvenTaps: :136-150 */
    case FT_FOLD_EVEN_ANTI:  /* ...AntiSymmetricEvenTaps: :151-165 (pre-subtract, exact) */
      for (int i = 0; i < N / 2; i += f->blk_sz, ram_addr += f->mww)
        for (int bc = f->blk_off, index = 0; bc < f->blk_off + f->blk_sz; bc++, index++) {
          w128 b = f->reg[N - 1 - i - index];
          w128 pre = f->ftype == FT_FOLD_EVEN ? f->reg[i + index] + b : f->reg[i + index] - b;
          acc = ob_macc(acc, &f->acc, pre * ram[ram_addr + bc], Fin + Fc);
        }
      break;
    case FT_FOLD_ODD:        /* ...SymmetricOddTaps: :166-185 (`fold` is ACC_TYPE, centre tap passes through) */
    case FT_FOLD_ODD_ANTI:   /* ...AntiSymmetricOddTaps: :186-205 */
      for (int i = 0; i < (N - 1) / 2 + 1; i += f->blk_sz, ram_addr += f->mww)
        for (int bc = f->blk_off, index = 0; bc < f->blk_off + f->blk_sz; bc++, index++) {
          w128 fold;
          if (i + index == (N - 1) / 2) fold = ob_convert(f->reg[i + index], Fin, &f->acc);
          else {
            w128 b = f->reg[N - 1 - i - index];
            fold = ob_convert(f->ftype == FT_FOLD_ODD ? f->reg[i + index] + b : f->reg[i + index] - b, Fin, &f->acc);
          }
          acc = ob_macc(acc, &f->acc, ram[ram_addr + bc] * fold, Fc + Fa);
        }
      break;
    default: return 0;       /* run() writes an unset core_out for the other FTYPE values (:277-303) */
  }
  return ob_convert(acc, Fa, &f->out);
}

/* n calls of run(data_in, coeffs, data_out) with the same coefficient RAM (raw values, n_ram words) */
long ob_rs_run(ob_rs *f, const int64_t *in, long n, const int64_t *ram_raw, int n_ram, int64_t *out) {
  if (!(f->ftype == FT_SHIFT_REG || f->ftype == FT_FOLD_EVEN || f->ftype == FT_FOLD_EVEN_ANTI || f->ftype == FT_FOLD_ODD ||
        f->ftype == FT_FOLD_ODD_ANTI)) return -1;
  if (n_ram < ob_rs_ram_words(f)) return -2;
  w128 *ram = (w128 *)calloc(n_ram > 0 ? n_ram : 1, sizeof(w128));
  for (int i = 0; i < n_ram; i++) ram[i] = ob_wrap((w128)ram_raw[i], f->coeff.W, f->coeff.S);
  for (long k = 0; k < n; k++) out[k] = (int64_t)rs_one(f, ob_wrap((w128)in[k], f->in.W, f->in.S), ram);
  free(ram);
  return n;
}
/* ac_firProgCoeffs_delay_line: :304-307 -> OUT_TYPE(reg[N_TAPS-1]) */
int64_t ob_rs_delay_out(ob_rs *f) { return (int64_t)ob_convert(f->reg[f->n - 1], F_of(&f->in), &f->out); }

/* ------------------------------------------------------------------ ac_poly_dec */
/* include/ac_dsp/ac_poly_dec.h:87-137: polyphase decimator.  taps[NTAPS*DF] shift register (zeroed, :95), DF partial
 * accumulators acc1[] and the total acc, all ACC_TYPE; coefficients in phase order coeffs[tp + NTAPS*df]. */
typedef struct {
  ob_fmt in, coeff, acc, out;
  int nt, df;
  w128 *taps, *h;
  w128 *pend; int npend;   /* samples of an incomplete group wait in data_in (`while (data_in.available(DF))`, :107) */
} ob_pd;

ob_pd *ob_pd_create(const ob_fmt *in, const ob_fmt *coeff, const ob_fmt *acc, const ob_fmt *out, int ntaps, int df) {
  ob_pd *f = (ob_pd *)calloc(1, sizeof(ob_pd));
  f->in = *in; f->coeff = *coeff; f->acc = *acc; f->out = *out; f->nt = ntaps; f->df = df;
  f->taps = (w128 *)calloc((size_t)ntaps * df, sizeof(w128));
  f->h = (w128 *)calloc((size_t)ntaps * df, sizeof(w128));
  f->pend = (w128 *)calloc((size_t)df, sizeof(w128));
  return f;
}
void ob_pd_destroy(ob_pd *f) { if (f) { free(f->taps); free(f->h); free(f->pend); free(f); } }
/* coeffs_t = coeffs_st.read(): :101-106 */
void ob_pd_load(ob_pd *f, const int64_t *c) {
  for (int i = 0; i < f->nt * f->df; i++) f->h[i] = ob_wrap((w128)c[i], f->coeff.W, f->coeff.S);
}
long ob_pd_run(ob_pd *f, const int64_t *in, long n, int64_t *out) {
  const int NT = f->nt, DF = f->df, L = NT * DF;
  const int Fin = F_of(&f->in), Fc = F_of(&f->coeff), Fa = F_of(&f->acc);
  long k = 0, nout = 0;
  while (f->npend + (n - k) >= DF) {                    /* a whole group of DF samples is available */
    w128 acc = 0;
    for (int df = DF - 1; df >= 0; df--) {              /* :110-123 */
      w128 x = f->npend > 0 ? f->pend[0] : ob_wrap((w128)in[k++], f->in.W, f->in.S);
      if (f->npend > 0) { memmove(f->pend, f->pend + 1, (size_t)(f->npend - 1) * sizeof(w128)); f->npend--; }
      for (int i = L - 1; i >= 0; i--) f->taps[i] = (i == 0) ? x : f->taps[i - 1];
      w128 acc1 = 0;                                    /* acc1[df] is 0 on entry (:96,122) */
      for (int tp = 0; tp < NT; tp++) acc1 = ob_macc(acc1, &f->acc, f->taps[tp * DF] * f->h[tp + NT * df], Fin + Fc);
      acc = ob_macc(acc, &f->acc, acc1, Fa);            /* acc = acc + acc1[df] */
    }
    out[nout++] = (int64_t)ob_convert(acc, Fa, &f->out);  /* OUT_TYPE acc_t = acc: :124-126 */
  }
  while (k < n) f->pend[f->npend++] = ob_wrap((w128)in[k++], f->in.W, f->in.S);
  return nout;
}

/* ------------------------------------------------------------------ ac_poly_intr */
/* include/ac_dsp/ac_poly_intr.h:103-256.  One step = one input sample -> IF outputs.  NTAPS is the length of the low-rate
 * delay line taps[] (:107,127-129).  Per phase j:
 *   FOLD_EVEN (:122-166): acc = sum_{i=NTAPS/2-1..0} coeffs[i + j*NTAPS/2] * fold_i, fold_i = ACC_TYPE(taps[i] + tp),
 *                         tp = sign[j] ? taps[NTAPS-1-i] : IN_TYPE(-taps[NTAPS-1-i])           (:137-146)
 *   FOLD_ODD  (:172-226): i = 0 .. (NTAPS-1)/2 upwards, centre tap fold = ACC_TYPE(taps[i]), coeffs[i + (NTAPS/2+1)*j]
 *   FOLD_ANTI (:232-251): plain polyphase MAC over taps[NTAPS-1..0] with coeffs[i + NTAPS*j], written at once.
 * The folded forms keep the accumulators of the PREVIOUS step in acc_a / acc_b (ping-pong on `flip`) and, from the
 * second step on (`init`), write per phase  t1 = prev[j]  or, when corr[j] != j (symmetric-pair coefficients),
 * (t1 + (sign[j] ? -prev[corr[j]] : prev[corr[j]])) >> 1  (:147-164): the outputs of step t are a function of step t-1. */
enum { PI_FOLD_EVEN, PI_FOLD_ODD, PI_FOLD_ANTI };
typedef struct {
  ob_fmt in, coeff, acc, out;
  int nt, ifac, ftype, csz, init;
  w128 *taps, *h, *prev;
  int *sign, *corr;
} ob_pi;

ob_pi *ob_pi_create(const ob_fmt *in, const ob_fmt *coeff, const ob_fmt *acc, const ob_fmt *out, int ntaps, int ifac, int ftype) {
  ob_pi *f = (ob_pi *)calloc(1, sizeof(ob_pi));
  f->in = *in; f->coeff = *coeff; f->acc = *acc; f->out = *out; f->nt = ntaps; f->ifac = ifac; f->ftype = ftype;
  f->csz = ifac * (ftype == PI_FOLD_EVEN ? ntaps / 2 : (ftype == PI_FOLD_ODD ? ntaps / 2 + 1 : ntaps));
  f->taps = (w128 *)calloc((size_t)ntaps, sizeof(w128));
  f->h = (w128 *)calloc((size_t)(f->csz > 0 ? f->csz : 1), sizeof(w128));
  f->prev = (w128 *)calloc((size_t)ifac, sizeof(w128));
  f->sign = (int *)calloc((size_t)ifac, sizeof(int));
  f->corr = (int *)calloc((size_t)ifac, sizeof(int));
  return f;
}
void ob_pi_destroy(ob_pi *f) { if (f) { free(f->taps); free(f->h); free(f->prev); free(f->sign); free(f->corr); free(f); } }
int ob_pi_coeffsz(const ob_pi *f) { return f->csz; }
/* ctrl_t = ctrl_st.read(); coeffs_t = coeffs_st.read(): :286-288 */
void ob_pi_load(ob_pi *f, const int64_t *c, const int64_t *sign, const int64_t *corr) {
  for (int i = 0; i < f->csz; i++) f->h[i] = ob_wrap((w128)c[i], f->coeff.W, f->coeff.S);
  for (int j = 0; j < f->ifac; j++) { f->sign[j] = sign[j] != 0; f->corr[j] = (int)(corr[j] & 0xff); }
}
/* -x assigned to a variable of format g: exact negation (one more bit), then the overflow mode of g */
static w128 pi_neg(w128 x, const ob_fmt *g) { return ob_convert(-x, F_of(g), g); }

long ob_pi_run(ob_pi *f, const int64_t *in, long n, int64_t *out) {
  const int NT = f->nt, IF = f->ifac;
  const int Fin = F_of(&f->in), Fc = F_of(&f->coeff), Fa = F_of(&f->acc);
  long nout = 0;
  w128 *cur = (w128 *)calloc((size_t)IF, sizeof(w128));
  for (long k = 0; k < n; k++) {
    for (int i = NT - 1; i >= 1; i--) f->taps[i] = f->taps[i - 1];
    f->taps[0] = ob_wrap((w128)in[k], f->in.W, f->in.S);
    for (int j = 0; j < IF; j++) {
      w128 acc = 0;
      if (f->ftype == PI_FOLD_EVEN) {
        for (int i = NT / 2 - 1; i >= 0; i--) {
          w128 tp = f->sign[j] ? f->taps[NT - 1 - i] : pi_neg(f->taps[NT - 1 - i], &f->in);
          w128 fold = ob_convert(f->taps[i] + tp, Fin, &f->acc);
          acc = ob_macc(acc, &f->acc, f->h[i + j * (NT / 2)] * fold, Fc + Fa);
        }
      } else if (f->ftype == PI_FOLD_ODD) {
        for (int i = 0; i < (NT - 1) / 2 + 1; i++) {
          w128 fold;
          if (i == (NT - 1) / 2) fold = ob_convert(f->taps[i], Fin, &f->acc);
          else {
            w128 tp = f->sign[j] ? f->taps[NT - 1 - i] : pi_neg(f->taps[NT - 1 - i], &f->in);
            fold = ob_convert(f->taps[i] + tp, Fin, &f->acc);
          }
          acc = ob_macc(acc, &f->acc, f->h[i + (NT / 2 + 1) * j] * fold, Fc + Fa);
        }
      } else {
        for (int i = NT - 1; i >= 0; i--) acc = ob_macc(acc, &f->acc, f->taps[i] * f->h[i + NT * j], Fin + Fc);
        out[nout++] = (int64_t)ob_convert(acc, Fa, &f->out);
        continue;
      }
      cur[j] = acc;
      if (f->init) {
        w128 t1 = f->prev[j];
        if (j != f->corr[j]) {
          w128 t2 = f->prev[f->corr[j]];
          w128 tn = f->sign[j] ? pi_neg(t2, &f->acc) : t2;
          out[nout++] = (int64_t)ob_convert((t1 + tn) >> 1, Fa, &f->out);   /* the sum has W_acc + 1 bits: >> 1 loses its LSB only */
        } else {
          out[nout++] = (int64_t)ob_convert(t1, Fa, &f->out);
        }
      }
    }
    if (f->ftype != PI_FOLD_ANTI) { memcpy(f->prev, cur, (size_t)IF * sizeof(w128)); f->init = 1; }
  }
  free(cur);
  return nout;
}

/* ------------------------------------------------------------------ ac_intg_dump */
/* include/ac_dsp/ac_intg_dump.h:84-151: integrate-and-dump over CHN interleaved channels.  Per frame one n_sample token
 * is read; samples j = 1 .. NS are added into temp[i] (ACC_TYPE, re-quantised at every add) and at j == n_sample the
 * CHN sums are dumped and cleared.  A token outside 1 .. NS never matches: the frame eats NS * CHN samples, dumps
 * nothing, and temp[] carries into the next frame (:133-147). */
typedef struct {
  ob_fmt in, acc, out;
  int ns, chn;
  w128 *temp;
} ob_id;

ob_id *ob_id_create(const ob_fmt *in, const ob_fmt *acc, const ob_fmt *out, int ns, int chn) {
  ob_id *f = (ob_id *)calloc(1, sizeof(ob_id));
  f->in = *in; f->acc = *acc; f->out = *out; f->ns = ns; f->chn = chn;
  f->temp = (w128 *)calloc((size_t)chn, sizeof(w128));
  return f;
}
void ob_id_destroy(ob_id *f) { if (f) { free(f->temp); free(f); } }

/* returns the number of outputs, or -1 if the samples do not cover the frames exactly */
long ob_id_run(ob_id *f, const int64_t *in, long n, const int64_t *nsamp, long nframes, int64_t *out) {
  const int Fin = F_of(&f->in), Fa = F_of(&f->acc);
  long k = 0, nout = 0;
  for (long fr = 0; fr < nframes; fr++) {
    const long long want = nsamp[fr];
    int flag = 0;
    for (int j = 1; j <= f->ns; j++) {                       /* ACC_LOOP :137 */
      for (int i = 0; i < f->chn; i++) {                     /* CHN_LOOP :138 */
        if (k >= n) return -1;
        w128 x = ob_wrap((w128)in[k++], f->in.W, f->in.S);
        f->temp[i] = ob_macc(f->temp[i], &f->acc, x, Fin);   /* temp[i] = temp[i] + data_in  :100 */
        if ((long long)j == want) {                          /* :101-105 */
          out[nout++] = (int64_t)ob_convert(f->temp[i], Fa, &f->out);
          f->temp[i] = 0;
          flag = 1;
        }
      }
      if (flag) break;
    }
  }
  return k == n ? nout : -1;
}

/* ------------------------------------------------------------------ ac_mv_avg -- PARITY UNPINNED */
/* include/ac_dsp/ac_mv_avg.h:99-127,154-196: weighted moving average over a window of TAPS samples centred on the point
 * being smoothed, burst by burst (n_sample samples per burst, the window restarts at every start-of-line flag):
 *   acc = 0;  for j = -TAPS/2 .. TAPS/2:  acc = acc + (ACC_TYPE) w[j] * coeffs[j + TAPS/2];  out = acc     (:111-121)
 * -- the sample is cast to ACC_TYPE BEFORE the multiply, the sum is re-quantised to ACC_TYPE at every tap.  w[j] comes
 * from ac_window_1d_flag (ac_math's ac_window.h, absent here): restated from the manual's description (section 2.4):
 * AC_CLIP repeats the edge sample, AC_MIRROR reflects about it, AC_WIN emits only the points whose window is full.
 * win: 0 = AC_WIN, 1 = AC_CLIP, 2 = AC_MIRROR.  n must be a multiple of n_sample, n_sample >= taps.  Returns #outputs. */
long ob_mvavg_run(const ob_fmt *in, const ob_fmt *out_f, const ob_fmt *acc, const ob_fmt *coeff, int taps, int win,
                  const int64_t *c, const int64_t *x, long n, long n_sample, int64_t *out) {
  const int Fin = F_of(in), Fa = F_of(acc), Fc = F_of(coeff), H = taps / 2;
  if (n_sample < taps || n_sample < 1 || n % n_sample) return -1;
  long nout = 0;
  for (long b = 0; b < n / n_sample; b++) {
    const int64_t *xb = x + b * n_sample;
    const long lo = win == 0 ? H : 0, hi = win == 0 ? n_sample - 1 - H : n_sample - 1;
    for (long i = lo; i <= hi; i++) {
      w128 a = 0;
      for (int j = -H; j <= H; j++) {
        long k = i + j;
        if (win == 1) { if (k < 0) k = 0; if (k > n_sample - 1) k = n_sample - 1; }
        else if (win == 2) { if (k < 0) k = -k; if (k > n_sample - 1) k = 2 * (n_sample - 1) - k; }
        const w128 s = ob_wrap((w128)xb[k], in->W, in->S);
        const w128 cast = ob_convert(s, Fin, acc);                                        /* (ACC_TYPE) w[j] */
        a = ob_macc(a, acc, cast * ob_wrap((w128)c[j + H], coeff->W, coeff->S), Fa + Fc);  /* acc_reg = acc_reg + ... */
      }
      out[nout++] = (int64_t)ob_convert(a, Fa, out_f);                                    /* data_out = acc_reg */
    }
  }
  return nout;
}

/* ------------------------------------------------------------------ CIC */
typedef struct {
  ob_fmt in, out, it;   /* it = lossless INT_TYPE */
  int R, M, N, intr;
  w128 *intg;           /* intg_reg[N]           ac_cic_full_core.h:74 */
  w128 *comb;           /* comb_dly_ln[N][M]     ac_cic_full_core.h:219 */
  unsigned rate_cnt, rate_cnt1, cnt;  /* 8-bit counters, ac_cic_full_core.h:72-73,91 */
  int dvalid;
  /* inf: samples handed from the low-rate to the high-rate half of intr run() */
  w128 *inf; long inf_head, inf_len, inf_cap;
} ob_cic;

static int log2_ceil_u(unsigned long long n) { int k = 0; while ((1ULL << k) < n) k++; return k; }

/* find_inter_type_cic_dec (ac_cic_dec_full.h:116-137) / _intr (ac_cic_intr_full.h:107-127) */
int ob_cic_int_width(int intr, const ob_fmt *in, int R, int M, int N) {
  unsigned long long g = 1;
  for (int i = 0; i < (intr ? N - 1 : N); i++) g *= (unsigned long long)R;
  for (int i = 0; i < N; i++) g *= (unsigned long long)M;
  return log2_ceil_u(g) + in->W + (in->S ? 0 : 1);
}

ob_cic *ob_cic_create(int intr, const ob_fmt *in, const ob_fmt *out, int R, int M, int N) {
  ob_cic *c = (ob_cic *)calloc(1, sizeof(ob_cic));
  c->in = *in; c->out = *out; c->R = R; c->M = M; c->N = N; c->intr = intr;
  c->it.W = ob_cic_int_width(intr, in, R, M, N);
  c->it.I = c->it.W - F_of(in);
  c->it.S = 1; c->it.Q = Q_TRN; c->it.O = O_WRAP;
  c->intg = (w128 *)calloc(N, sizeof(w128));
  c->comb = (w128 *)calloc((size_t)N * M, sizeof(w128));
  /* ac_cic_full_core_intg ctor, ac_cic_full_core.h:94-102 (both classes pass value=true) */
  c->rate_cnt = 0; c->dvalid = 1; c->cnt = 0; c->rate_cnt1 = (unsigned)(R - 1) & 0xFFu;
  return c;
}
void ob_cic_destroy(ob_cic *c) { if (c) { free(c->intg); free(c->comb); free(c->inf); free(c); } }

/* intStage: ac_cic_full_core.h:80-87 */
static w128 cic_int_stage(ob_cic *c, w128 x) {
  const int Fi = F_of(&c->it);
  for (int i = c->N - 1; i > 0; i--) c->intg[i] = ob_convert(c->intg[i] + c->intg[i - 1], Fi, &c->it);
  c->intg[0] = ob_convert(x + c->intg[0], Fi, &c->it);
  return c->intg[c->N - 1];
}
/* comb + diffStage: ac_cic_full_core.h:228-255.  The delay line is shifted with an ASCENDING copy loop
 * (`for i = 0 .. M-1: if (i != 0) dly[i] = dly[i-1]`, :247-251), so dly[0] smears through the whole line: after a step
 * every dly[i >= 1] holds the previous input and dly[M-1] read at the next step is the input of two steps ago.  The
 * differential delay of the reference as written is therefore min(M, 2), not M, while the lossless width
 * (find_inter_type_cic_*) still grows with M.  Found by the random-instantiation sweep (tests/test_oracle_fuzz.py):
 * the reference's own vectors use M = 2, where both readings coincide.  Restated literally. */
static w128 cic_comb(ob_cic *c, w128 x) {
  const int Fi = F_of(&c->it);
  w128 v = x;
  for (int k = 0; k < c->N; k++) {
    w128 *d = c->comb + (size_t)k * c->M;
    w128 o = ob_convert(v - d[c->M - 1], Fi, &c->it);
    for (int i = 1; i < c->M; i++) d[i] = d[i - 1];
    d[0] = v;
    v = o;
  }
  return v;
}

static void inf_push(ob_cic *c, w128 v) {
  if (c->inf_head + c->inf_len == c->inf_cap) {
    if (c->inf_head > 0) { memmove(c->inf, c->inf + c->inf_head, (size_t)c->inf_len * sizeof(w128)); c->inf_head = 0; }
    else { c->inf_cap = c->inf_cap ? 2 * c->inf_cap : 1024; c->inf = (w128 *)realloc(c->inf, (size_t)c->inf_cap * sizeof(w128)); }
  }
  c->inf[c->inf_head + c->inf_len++] = v;
}

/* run(): ac_cic_dec_full.h:163-222 (decIntg -> inf -> decDiff) or
 *        ac_cic_intr_full.h:150-215 (intrDiff -> inf -> intrIntg). Returns outputs written. */
long ob_cic_run(ob_cic *c, const int64_t *in, long n, int64_t *out) {
  const int Fi = F_of(&c->it);
  long k = 0;
  if (!c->intr) {
    for (long j = 0; j < n; j++) {
      /* decIntgCore: ac_cic_full_core.h:110-135 */
      w128 x = ob_convert(ob_wrap((w128)in[j], c->in.W, c->in.S), F_of(&c->in), &c->it);
      int valid = (c->rate_cnt == 0);
      w128 y = cic_int_stage(c, x);
      c->dvalid = valid;
      c->rate_cnt = (c->rate_cnt + 1) & 0xFFu;
      if (c->rate_cnt > (unsigned)(c->R - 1)) c->rate_cnt = 0;
      /* decIntg writes inf when dvalid (ac_cic_dec_full.h:196-198); decDiff drains it (:213-221) */
      if (c->dvalid) out[k++] = (int64_t)ob_convert(cic_comb(c, y), Fi, &c->out);
    }
    return k;
  }
  /* intrDiff: ac_cic_intr_full.h:173-185 */
  for (long j = 0; j < n; j++) {
    w128 x = ob_convert(ob_wrap((w128)in[j], c->in.W, c->in.S), F_of(&c->in), &c->it);
    inf_push(c, cic_comb(c, x));
  }
  /* intrIntg: ac_cic_intr_full.h:195-215; data_in_t is a local re-initialised to 0 per call (:196) */
  w128 data_in_t = 0;
  while (c->inf_len > 0) {
    if (c->dvalid) { data_in_t = c->inf[c->inf_head++]; c->inf_len--; }
    /* intrIntgCore: ac_cic_full_core.h:143-160 */
    w128 feed;
    if (c->rate_cnt1 == ((unsigned)(c->R - 1) & 0xFFu)) { feed = data_in_t; c->rate_cnt1 = 0; c->dvalid = 0; }
    else if (c->rate_cnt1 == ((unsigned)(c->R - 2) & 0xFFu)) { feed = 0; c->rate_cnt1 = (c->rate_cnt1 + 1) & 0xFFu; c->dvalid = 1; }
    else { feed = 0; c->rate_cnt1 = (c->rate_cnt1 + 1) & 0xFFu; c->dvalid = 0; }
    w128 y = cic_int_stage(c, feed);
    if (c->cnt < (unsigned)(c->N - 1)) c->cnt = (c->cnt + 1) & 0xFFu;
    else out[k++] = (int64_t)ob_convert(y, Fi, &c->out);
  }
  if (c->inf_len == 0) c->inf_head = 0;
  return k;
}
