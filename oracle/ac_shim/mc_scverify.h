// oracle/ac_shim/mc_scverify.h -- TEST INFRASTRUCTURE. Standalone-compile stub:
// the reference wraps its top-level run() in CCS_BLOCK(); outside Catapult it is the identity.
#ifndef B200DSP_ORACLE_AC_SHIM_MC_SCVERIFY_H
#define B200DSP_ORACLE_AC_SHIM_MC_SCVERIFY_H
#define CCS_BLOCK(a) a
#define CCS_MAIN(a, b) int main(a, b)
#define CCS_RETURN(a) return (a)
#endif
