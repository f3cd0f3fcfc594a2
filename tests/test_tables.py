"""SG4 integer tables: oracle vs the reference's golden logs, product (C-ABI) vs oracle, bit-exact."""
import numpy as np
import pytest

from oracle.sg4_oracle import Tables

CASES = {   # name -> (D, LB, LG, A, B, legacy_LB0)
    "HNO3_LB0_LG3": (8, 0, 3, 1, 1, True),     # log written by ElVibRot 181.3 (nb(L) uncapped at LB=0)
    "HNO3_LB1_LG3": (8, 1, 3, 1, 1, False),
    "HNO3_LB2_LG4": (8, 2, 4, 1, 1, False),
    "HNO3_LB3_LG5": (8, 3, 5, 1, 1, False),
    "HCN_LB6_LG7": (3, 6, 7, [10, 1, 1], [10, 2, 2], False),
}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_tables_match_reference_logs(name, golden):
    D, LB, LG, A, B, legacy = CASES[name]
    g = golden["sg4_tables"][name]
    t = Tables(D, LB, LG, A, B, legacy_LB0=legacy)
    assert t.Lmin == g["Lmin"] and LG == g["LG"]
    assert t.nb_SG == g["nb_SG"]
    assert t.nb == g["nb"]
    assert t.S == g["S"] == g["nbb"]
    assert t.NQ == g["nqq"]
    assert t.count0 == g["count0"] == int((t.map == 0).sum())
    assert int(t.tab_nq.max()) == g["max_nq"] and int(t.tab_nb.max()) == g["max_nb"]
    for k, row in g["nb_of"].items():
        assert list(t.nb_of[int(k) - 1]) == row
    for k, row in g["nq_of"].items():
        assert list(t.nq_of[int(k) - 1]) == row
    for term in g["terms"]:                      # full (iG, l, weight) table where the log prints it
        i = term["iG"] - 1
        assert list(t.tab_l[i]) == term["l"]
        assert t.weight[i] == term["w"]
    for p in g["packed_first"]:                  # first 100 packed multi-indices
        assert list(t.packedB[p["ib"] - 1]) == p["idx"]
    # every packed function is reached by at least one term (reference check :908-915)
    assert (np.bincount(t.map, minlength=t.nb + 1)[1:] > 0).all()


def test_current_source_LB0_caps_nb():
    """ElVibRot 184.1 (the mounted source) does not extrapolate a one-entry L_TO_nb table
    (sub_module_Basis_LTO_n.f90:431-440): with LB=0 every level has nb=1."""
    t = Tables(8, 0, 3, 1, 1)
    assert (t.nb_of == 1).all() and t.nb == 1 and t.S == t.nb_SG == 165 and t.count0 == 0


PRODUCT_CASES = [(8, 2, 4, 1, 1), (8, 3, 5, 1, 1), (3, 6, 7, [10, 1, 1], [10, 2, 2]), (3, 4, 5, [10, 1, 1], [10, 2, 2]),
                 (6, 3, 3, 1, 2), (21, 2, 2, 1, 2), (12, 4, 4, 1, 2), (12, 1, 1, 1, [3, 3] + [2] * 10),
                 (5, 2, 6, 2, 3), (1, 3, 3, 1, 2), (2, 0, 2, 1, 1), (4, 5, 3, 1, 2)]


@pytest.mark.parametrize("case", PRODUCT_CASES, ids=[str(c[:3]) for c in PRODUCT_CASES])
def test_product_tables_bit_exact_with_oracle(case, evr):
    D, LB, LG, A, B = case
    b = evr.workloads.hm_sg4_basis(D, LB, LG, A, B)
    t = Tables(D, LB, LG, A, B)
    assert (b.nb_SG, b.nb, b.Max_Srep, b.nqq, b.count0, b.Lmin) == (t.nb_SG, t.nb, t.S, t.NQ, t.count0, t.Lmin)
    assert np.array_equal(b.nq_of, t.nq_of) and np.array_equal(b.nb_of, t.nb_of)
    assert np.array_equal(b.nDind_SmolyakRep_Tab_nDval, t.tab_l)
    assert np.array_equal(b.WeightSG, t.weight)
    assert np.array_equal(b.tab_nq_OF_SRep, t.tab_nq) and np.array_equal(b.tab_nb_OF_SRep, t.tab_nb)
    assert np.array_equal(b.tab_Sum_nq_OF_SRep, t.sum_nq) and np.array_equal(b.tab_Sum_nb_OF_SRep, t.sum_nb)
    assert np.array_equal(b.nDindB_Tab_nDval, t.packedB)
    assert np.array_equal(b.tab_iB_OF_SRep_TO_iB, t.map)


def test_hh12d_sweep_sizes(evr):
    """Sizes of the throughput configuration."""
    expect = {2: (91, 313, 691), 3: (455, 2625, 8695), 4: (1820, 16641, 83020), 5: (6188, 85305, 642172)}
    for L, (nsg, nb, nq) in expect.items():
        b = evr.workloads.hm_sg4_basis(12, L, L, 1, 2)
        assert (b.nb_SG, b.nb, b.nqq) == (nsg, nb, nq)


def test_ini_iGs_matches_reference_formula(evr):
    import ctypes as C
    L = evr.lib.lib()
    for nb_SG, np_ in [(85, 2), (85, 3), (13, 2), (50388, 8), (5, 8), (7, 1)]:
        q, rem = divmod(nb_SG, np_)
        prev = 0
        for r in range(np_):
            b, e = C.c_int(), C.c_int()
            assert L.evr_sg4_ini_iGs(nb_SG, np_, r, C.byref(b), C.byref(e)) == 0
            b1 = r * q + 1 + min(r, rem)                        # ini_iGs_MPI, 1-based inclusive
            b2 = (r + 1) * q + min(r, rem) + (1 if rem > r else 0)
            assert (b.value, e.value) == (b1 - 1, b2)
            assert b.value == prev
            prev = e.value
        assert prev == nb_SG


def test_balanced_iGs_partitions_by_cost(evr):
    b = evr.workloads.hm_sg4_basis(12, 4, 4, 1, 2)
    cost = b.tab_nq_OF_SRep
    for np_ in (1, 2, 3, 8):
        prev, loads = 0, []
        for r in range(np_):
            lo, hi = evr.distributed.balanced_iGs(cost, np_, r)
            assert lo == prev and hi >= lo
            loads.append(int(cost[lo:hi].sum()))
            prev = hi
        assert prev == b.nb_SG and sum(loads) == b.nqq
        assert max(loads) - min(loads) <= 2 * int(cost.max())


@pytest.mark.parametrize("L", [6, 7])
def test_product_tables_bit_exact_with_oracle_benchmarked_sizes(L, evr):
    """The benchmarked configuration (HH 12-D, LB = LG = 6 and 7: 18 564 / 50 388 terms, 4.2 M / 23.8 M entries of
    tab_iB_OF_SRep_TO_iB): the product's closed-form tables against the oracle's enumeration + search, bit-exact."""
    b = evr.workloads.hm_sg4_basis(12, L, L, 1, 2)
    t = Tables(12, L, L, 1, 2)
    expect = {6: (18564, 369305, 4195284), 7: (50388, 1392065, 23826372)}[L]
    assert (b.nb_SG, b.nb, b.nqq) == expect == (t.nb_SG, t.nb, t.NQ)
    assert (b.Max_Srep, b.count0, b.Lmin) == (t.S, t.count0, t.Lmin)
    assert np.array_equal(b.nDind_SmolyakRep_Tab_nDval, t.tab_l)
    assert np.array_equal(b.WeightSG, t.weight)
    assert np.array_equal(b.tab_nq_OF_SRep, t.tab_nq) and np.array_equal(b.tab_nb_OF_SRep, t.tab_nb)
    assert np.array_equal(b.tab_Sum_nq_OF_SRep, t.sum_nq) and np.array_equal(b.tab_Sum_nb_OF_SRep, t.sum_nb)
    assert np.array_equal(b.nDindB_Tab_nDval, t.packedB)
    assert np.array_equal(b.tab_iB_OF_SRep_TO_iB, t.map)
