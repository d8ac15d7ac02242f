// oracle/ac_shim/ac_window.h -- TEST INFRASTRUCTURE.  PARITY UNPINNED.
//
// ac_window.h belongs to hlslibs/ac_math, which the reference does not vendor and this image does not contain; the
// reference ships no test or vector for ac_mv_avg either.  This file restates the ONE class ac_mv_avg.h uses --
// ac_window_1d_flag<T, AC_WN, AC_WMODE> with write(value, sol, eol), valid() and operator[] -- from the behaviour the
// ac_dsp manual documents for the moving-average block (pdfdocs/ac_dsp_ref.pdf section 2.4: a window of AC_WN samples
// centred on the point being smoothed; at the ends of a burst either clipping -- the edge sample repeated -- or mirroring
// about the edge sample; "no boundary processing" otherwise) and from the way ac_mv_avg::run() drives it
// (include/ac_dsp/ac_mv_avg.h:154-196: sol on the first sample of a burst, eol on the last, AC_WN/2 more writes after eol
// in the boundary modes so that the window can flush).  Edge cases the manual does not cover (bursts shorter than the
// window) are rejected by the engine rather than guessed.  Nothing here is copied from ac_math.
#ifndef _INCLUDED_AC_WINDOW_H_SHIM_
#define _INCLUDED_AC_WINDOW_H_SHIM_

#include <vector>

enum ac_window_mode { AC_WIN = 0, AC_CLIP = 1, AC_MIRROR = 2 };

template <class T, int AC_WN, int AC_WMODE = AC_WIN>
class ac_window_1d_flag {
public:
  ac_window_1d_flag() : k_(-1), n_(-1), valid_(false) {}

  // One sample enters the window.  sol: first sample of a burst (the window restarts); eol: last sample of the burst.
  void write(T src, bool sol, bool eol) {
    if (sol) { burst_.clear(); k_ = -1; n_ = -1; }
    k_++;
    if (n_ < 0) burst_.push_back(src);             // writes after eol only push the window along (their value is not used)
    if (eol) n_ = k_ + 1;
    const int c = k_ - AC_WN / 2;                  // index of the sample at the centre of the window
    if (AC_WMODE == AC_WIN) valid_ = n_ < 0 ? (k_ >= AC_WN - 1) : (k_ >= AC_WN - 1 && k_ < n_);
    else valid_ = c >= 0 && (n_ < 0 || c < n_);
  }
  bool valid() const { return valid_; }

  // w[j], j = -AC_WN/2 .. AC_WN/2: the sample j places from the centre, boundary rule applied
  const T &operator[](int j) const {
    int i = k_ - AC_WN / 2 + j;
    const int last = (n_ < 0 ? k_ : n_ - 1);      // newest real sample of the burst
    if (AC_WMODE == AC_CLIP) { if (i < 0) i = 0; if (i > last) i = last; }
    else if (AC_WMODE == AC_MIRROR) { if (i < 0) i = -i; if (i > last) i = 2 * last - i; }
    if (i < 0) i = 0;
    if (i > last) i = last;
    return burst_[(size_t)i];
  }

private:
  std::vector<T> burst_;
  int k_, n_;
  bool valid_;
};

#endif
