#!/usr/bin/env python
"""Extract golden vectors for the SG4 H|psi> path from the reference tree.

Run in the build container (the reference is mounted read-only at /root/reference;
it does not exist on the GPU box, which only sees the JSON written here):

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/*.json.  Only numbers printed by the reference's
own regression logs / benchmark files / quadrature tables are extracted -- no source.
"""
import gzip
import json
import os
import re
import sys

REF = os.environ.get("EVR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _ints(s):
    return [int(x) for x in s.split()]


def parse_sg4_log(path):
    """Pull the SG4 table printout out of a vib log (RecSparseGrid_ForDP_type4 /
    Set_tables_FOR_SmolyakRepBasis_TO_tabPackedBasis prints)."""
    with gzip.open(path, "rt", errors="replace") as f:
        lines = f.read().splitlines()
    g = {"source": os.path.relpath(path, REF), "nb_of": {}, "nq_of": {}, "terms": [], "packed_first": []}
    in_sg4 = False
    for ln in lines:
        if "SPARSE GRID type4" in ln:
            in_sg4 = True
        m = re.match(r"\s*- Sparse Grid, Lmin,Lmax:\s+(\d+)\s+(\d+)", ln)
        if m:
            g["Lmin"], g["LG"] = int(m.group(1)), int(m.group(2))
        m = re.match(r"\s*(\d+) nb\(L\)\s+(.*)$", ln)
        if m and in_sg4:
            g["nb_of"][m.group(1)] = _ints(m.group(2))
        m = re.match(r"\s*(\d+) nq\(L\)\s+(.*)$", ln)
        if m and in_sg4:
            g["nq_of"][m.group(1)] = _ints(m.group(2))
        m = re.match(r"\s*i_SG,nDval,coef\s+(.*)$", ln)
        if m and "...." not in ln:
            t = m.group(1).split()
            g["terms"].append({"iG": int(t[0]), "l": [int(x) for x in t[1:-1]], "w": float(t[-1])})
        m = re.match(r"\s*ib,tab_L\s+(\d.*)$", ln)
        if m:
            t = _ints(m.group(1))
            g["packed_first"].append({"ib": t[0], "idx": t[1:]})
        for key, pat in [("nb_SG", r"\s*nb of terms \(grids\)\s+(\d+)"), ("nb", r"\s*nb_ba\s+(\d+)\s*$"),
                         ("S", r"\s*Max_Srep\s+(\d+)"), ("count0", r"\s*count 0\s+(\d+)"),
                         ("nbb", r"\s*nbb\s+\(Smolyak Rep\)\s+(\d+)"), ("nqq", r"\s*nqq\s+\(Smolyak Rep\)\s+(\d+)")]:
            m = re.match(pat, ln)
            if m and key not in g:
                g[key] = int(m.group(1))
        m = re.match(r"\s*max nq nb:\s+(\d+)\s+(\d+)", ln)
        if m and "max_nq" not in g:
            g["max_nq"], g["max_nb"] = int(m.group(1)), int(m.group(2))
        m = re.match(r"\s*Working with (\S+)", ln)
        if m:
            g["version"] = m.group(1)
    return g


def main():
    if not os.path.isdir(REF):
        sys.exit(f"reference not found at {REF}")
    gold = {}
    # --- integer-table golden logs -------------------------------------------------
    logs = {
        "HNO3_LB0_LG3": "UnitTests/HNO3_UT/RES_old/res_HNO3_RPH_LB0-LG3.gz",
        "HNO3_LB1_LG3": "UnitTests/HNO3_UT/RES_old/res_HNO3_RPH_LB1-LG3.gz",
        "HNO3_LB2_LG4": "UnitTests/HNO3_UT/RES_old/res_HNO3_RPH_LB2-LG4.gz",
        "HNO3_LB3_LG5": "UnitTests/HNO3_UT/RES_old/res_HNO3_RPH_LB3-LG5.gz",
        "HCN_LB6_LG7": "UnitTests/HCN_UT/RES_old/res_RPH_AutoContract_Davidson_SG4.gz",
    }
    for k, rel in logs.items():
        gold[k] = parse_sg4_log(os.path.join(REF, rel))
    with open(os.path.join(OUT, "sg4_tables.json"), "w") as f:
        json.dump(gold, f, indent=0, separators=(",", ":"))
    # --- eigenvalue / autocorrelation known answers -----------------------------------
    kat = {}
    for k, rel in [("HH6D_L3", "Working_tests/MPI_tests/6D_Davidson_openMP/benchmark"),
                   ("HH21D_L2", "Working_tests/MPI_tests/21D_Davidson_openMP/benchmark")]:
        with open(os.path.join(REF, rel)) as f:
            rows = [[float(x) for x in ln.split()] for ln in f if ln.strip()]
        kat[k] = {"source": rel, "unit": "au", "tol": 1e-8, "levels": [r[0] for r in rows]}
    for k, rel in [("PYR12D_L1_autocor", "Working_tests/MPI_tests/12D_propagation_openMP/benchmark")]:
        with open(os.path.join(REF, rel)) as f:
            rows = [[float(x) for x in ln.split()] for ln in f if ln.strip()]
        kat[k] = {"source": rel, "tol": 1e-8, "t_re_im_abs": rows}
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=0, separators=(",", ":"))
    # --- Gauss-Hermite tables read by the Hm basis (sub_quadra_herm.f90:328-348) --------
    herm = {}
    for nq in range(1, 41):
        p = os.path.join(REF, f"Internal_data/HermQuadra/herm{nq}.txt")
        xs, ws = [], []
        with open(p) as f:
            for ln in f:
                t = ln.split()
                if len(t) >= 4:
                    xs.append(float(t[1]))
                    ws.append(float(t[3]))      # 4th column = w * exp(x^2)
        assert len(xs) == nq, (nq, len(xs))
        herm[str(nq)] = {"x": xs, "w": ws}
    with open(os.path.join(OUT, "herm_quadra.json"), "w") as f:
        json.dump(herm, f, separators=(",", ":"))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
