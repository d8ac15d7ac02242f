// oracle/ref_driver_cic.h -- TEST INFRASTRUCTURE. Shared base for the two CIC driver TUs.
#ifndef B200DSP_ORACLE_REF_DRIVER_CIC_H
#define B200DSP_ORACLE_REF_DRIVER_CIC_H
namespace acref {
struct CicBase {
  double run_seconds;  // time inside the reference's run() during the last call (channel fill / drain excluded)
  CicBase() : run_seconds(0) {}
  virtual ~CicBase() {}
  virtual long run(const long long *in, long n, long long *out) = 0;
};
}  // namespace acref
#endif
