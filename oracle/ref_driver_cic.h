// oracle/ref_driver_cic.h -- TEST INFRASTRUCTURE. Shared base for the two CIC driver TUs.
#ifndef B200DSP_ORACLE_REF_DRIVER_CIC_H
#define B200DSP_ORACLE_REF_DRIVER_CIC_H
namespace acref {
struct CicBase {
  virtual ~CicBase() {}
  virtual long run(const long long *in, long n, long long *out) = 0;
};
}  // namespace acref
#endif
