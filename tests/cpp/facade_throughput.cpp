// tests/cpp/facade_throughput.cpp -- what the reference-signature calls cost: samples per second through
// ac_channel-based run(), for the block classes (everything queued, one run()) and for ac_fir_prog_coeffs (ONE sample per
// run() call, reference include/ac_dsp/ac_fir_prog_coeffs.h:277-303).  The SAME source compiles against the facade
// (-I include/b200dsp: the B200 engine) and against the reference headers (-I <reference>/include: the CPU templates);
// tools/facade_throughput.sh builds both and prints the two sets of lines side by side.
//
//   facade_throughput [n_block] [n_prog]
#include <ac_fixed.h>
#include <ac_channel.h>
#include <ac_dsp/ac_fir_load_coeffs.h>
#include <ac_dsp/ac_fir_prog_coeffs.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifndef IMPL
#define IMPL "facade"
#endif

typedef ac_fixed<16, 1, true> IN_T;
typedef ac_fixed<40, 8, true> ACC_T;
const unsigned N_TAPS = 256;

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F, class Drive>
static void time_case(const char *name, size_t n, Drive drive) {
  drive(n > 4096 ? 4096 : n);                       // warm-up: creates the engine handle, loads the taps
  const double t0 = now();
  const double sum = drive(n);
  const double dt = now() - t0;
  std::printf("{\"impl\": \"%s\", \"case\": \"%s\", \"samples\": %zu, \"seconds\": %.6f, \"samples_per_s\": %.1f, \"checksum\": %.0f}\n", IMPL, name, n, dt,
              n / dt, sum);
  std::fflush(stdout);
}

int main(int argc, char **argv) {
  const size_t n_block = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)1 << 20;
  const size_t n_prog = argc > 2 ? (size_t)atoll(argv[2]) : 4096;
  std::vector<IN_T> taps(N_TAPS);
  unsigned s = 12345;
  for (unsigned i = 0; i < N_TAPS; i++) { s = s * 1664525u + 1013904223u; taps[i].set_slc(0, ac_int<16, true>((int)((s >> 16) & 0xFFFF) - 32768)); }

  {  // ac_fir_load_coeffs: taps through the coefficient channel, then every sample queued and ONE run()
    static ac_fir_load_coeffs<IN_T, ACC_T, IN_T, ACC_T, N_TAPS, SHIFT_REG> f;
    ac_channel<IN_T> in, co;
    ac_channel<ACC_T> out;
    ac_channel<bool> ld;
    for (unsigned i = 0; i < N_TAPS; i++) co.write(taps[i]);
    ld.write(true);
    f.run(in, co, out, ld);
    time_case<int>("ac_fir_load_coeffs 256 taps: all samples queued, one run()", n_block, [&](size_t n) {
      unsigned r = 777;
      for (size_t i = 0; i < n; i++) { r = r * 1664525u + 1013904223u; IN_T v; v.set_slc(0, ac_int<16, true>((int)(r >> 16) - 32768)); in.write(v); }
      f.run(in, co, out, ld);
      double sum = 0;
      while (out.available(1)) sum += out.read().to_double();
      return sum;
    });
  }
  {  // ac_fir_prog_coeffs: the reference's calling convention, one sample per run()
    static ac_fir_prog_coeffs<IN_T, ACC_T, IN_T, ACC_T, (int)N_TAPS, SHIFT_REG> f;
    ac_channel<IN_T> in;
    ac_channel<ACC_T> out;
    time_case<int>("ac_fir_prog_coeffs 256 taps: one sample per run() call", n_prog, [&](size_t n) {
      unsigned r = 999;
      double sum = 0;
      for (size_t i = 0; i < n; i++) {
        r = r * 1664525u + 1013904223u;
        IN_T v; v.set_slc(0, ac_int<16, true>((int)(r >> 16) - 32768));
        in.write(v);
        f.run(in, out, taps.data());
        sum += out.read().to_double();
      }
      return sum;
    });
#ifdef B200DSP_FIR_BLOCK_H
    // facade extension: the same object fed a block at a time (== calling run() until the channel is empty)
    time_case<int>("ac_fir_prog_coeffs 256 taps: run_block() extension", n_block, [&](size_t n) {
      unsigned r = 555;
      for (size_t i = 0; i < n; i++) { r = r * 1664525u + 1013904223u; IN_T v; v.set_slc(0, ac_int<16, true>((int)(r >> 16) - 32768)); in.write(v); }
      f.run_block(in, out, taps.data());
      double sum = 0;
      while (out.available(1)) sum += out.read().to_double();
      return sum;
    });
#endif
  }
  return 0;
}
