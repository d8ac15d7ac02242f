"""oracle/gen_ref_cfgs.py -- TEST INFRASTRUCTURE.
Emit the X-macro lists (oracle/_ref/cfgs_*.inc) that ref_driver_*.cpp include."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_configs as rc  # noqa: E402


def f(t):
    W, I, S, Q, O = t
    return f"{W},{I},{'true' if S else 'false'},{Q},{O}"


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "cfgs_fir.inc"), "w") as fh:
        for cid, name, fi, fc, fa, fo, t in rc.fir_configs():
            fh.write(f"X({cid}, {f(fi)}, {f(fc)}, {f(fa)}, {f(fo)}, {t})\n")
    with open(os.path.join(outdir, "cfgs_rs.inc"), "w") as fh:
        for cid, (N, fi, fo, fc, fa, mww, bs, bo, ft) in enumerate(rc.RS_CONFIGS):
            fh.write(f"X({cid}, {N}, {f(fi)}, {f(fo)}, {f(fc)}, {f(fa)}, {mww}, {bs}, {bo}, {ft}, {rc.rs_ram_words(rc.RS_CONFIGS[cid])})\n")
    with open(os.path.join(outdir, "cfgs_pd.inc"), "w") as fh:
        for cid, (fi, fc, fa, fo, nt, df) in enumerate(rc.PD_CONFIGS):
            fh.write(f"X({cid}, {f(fi)}, {f(fc)}, {f(fa)}, {f(fo)}, {nt}, {df})\n")
    with open(os.path.join(outdir, "cfgs_pi.inc"), "w") as fh:
        for cid, cfg in enumerate(rc.PI_CONFIGS):
            fi, fc, fa, fo, nt, IF, ft = cfg
            fh.write(f"X({cid}, {f(fi)}, {f(fc)}, {f(fa)}, {f(fo)}, {nt}, {rc.pi_coeffsz(cfg)}, {IF}, {ft})\n")
    with open(os.path.join(outdir, "cfgs_id.inc"), "w") as fh:
        for cid, (fi, fa, fo, ns, chn) in enumerate(rc.ID_CONFIGS):
            fh.write(f"X({cid}, {f(fi)}, {f(fa)}, {f(fo)}, {ns}, {chn})\n")
    with open(os.path.join(outdir, "cfgs_mv.inc"), "w") as fh:
        for cid, (maxs, taps, wt, fi, fo, fa, fc) in enumerate(rc.MV_CONFIGS):
            fh.write(f"X({cid}, {maxs}, {taps}, {wt}, {f(fi)}, {f(fo)}, {f(fa)}, {f(fc)})\n")
    for mode in ("dec", "intr"):
        with open(os.path.join(outdir, f"cfgs_cic_{mode}.inc"), "w") as fh:
            for cid, c in enumerate(rc.CIC_CONFIGS):
                if c[0] != mode:
                    continue
                _, R, M, N, fi, fo = c
                fh.write(f"X({cid}, {R}, {M}, {N}, {f(fi)}, {f(fo)})\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))
