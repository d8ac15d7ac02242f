// oracle/ac_shim/ac_channel.h -- TEST INFRASTRUCTURE, not product code.
// Clean-room stand-in for AC Datatypes `ac_channel<T>`: an unbounded FIFO with the
// calls the ac_dsp headers and test benches make (read, write, available, debug_size).
#ifndef B200DSP_ORACLE_AC_SHIM_AC_CHANNEL_H
#define B200DSP_ORACLE_AC_SHIM_AC_CHANNEL_H

#include <deque>
#include <cstdlib>
#include <cstdio>

template <class T>
class ac_channel {
  std::deque<T> q;
public:
  ac_channel() {}
  T read() {
    if (q.empty()) { std::fprintf(stderr, "ac_channel: read from empty channel\n"); std::abort(); }
    T t = q.front();
    q.pop_front();
    return t;
  }
  void read(T &t) { t = read(); }
  void write(const T &t) { q.push_back(t); }
  bool available(unsigned n) const { return q.size() >= n; }
  unsigned debug_size() const { return (unsigned)q.size(); }
  unsigned size() const { return (unsigned)q.size(); }
  bool nb_read(T &t) { if (q.empty()) return false; t = read(); return true; }
};

#endif
