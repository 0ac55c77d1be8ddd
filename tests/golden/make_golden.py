#!/usr/bin/env python
"""Copies the reference's known-answer DATA fixtures (CSV histograms and learned matrices,
/root/reference/test/data/*.csv, checked by test/runtests.jl:68-101) into tests/golden/.
The reference (Julia + Ipopt) cannot run in the build container, so these stored files are the
only outputs of the real reference available; /root/reference does not exist on the GPU box,
hence the copy.  Run once:  python tests/golden/make_golden.py
"""
import pathlib
import shutil

SRC = pathlib.Path("/root/reference/test/data")
DST = pathlib.Path(__file__).resolve().parent

if __name__ == "__main__":
    for name in ("a", "b", "c", "mvt"):
        shutil.copy(SRC / f"{name}_samples.csv", DST / f"{name}_samples.csv")
        for form in ("RISE", "logRISE", "RPLE"):
            shutil.copy(SRC / f"{name}_{form}_learned.csv", DST / f"{name}_{form}_learned.csv")
    print("copied", len(list(DST.glob("*.csv"))), "csv fixtures")
