#!/usr/bin/env python
"""Regenerates the committed fixtures in this directory.

reference_pins.json -- the numeric known-answers the REFERENCE's own tests / doctests hold for the vector-query
path (all flat-path), copied as data with their source locations.  They pin the
oracle (tests/test_oracle_pins.py) and, through it, the CUDA path.

ivfpq_small.npz -- a small IVF_PQ index (plain arrays), queries, and the ORACLE's answers for them (l2 /
cosine / dot; plain, range-filtered, prefiltered, refined).  The reference cannot be run here (Rust, no
toolchain), so this fixture is oracle output, not reference output: it freezes the restatement between rounds
(a change to oracle.c that alters any value fails `-m "not gpu"`), and the GPU tests check the CUDA path
against the same committed bytes.  "IVF_PQ parity unpinned" (oracle/oracle.h) still applies.

lance_ivfpq_small/ -- the cosine index of ivfpq_small.npz written as Lance index files (index.idx + auxiliary.idx)
by lancedb_b200/lance_index.py's fixture writer.  The layout is RECALLED, not verified against a real Lance file
(none exists in the reference tree): the fixture freezes the reader's and writer's bytes between rounds
(tests/test_host_logic.py), nothing more.

Run from the repo root:  python tests/golden/make_golden.py   (does not rewrite ivfpq_small.npz unless --all)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

PINS = {
    "_comment": "known-answer values held by the reference's own tests/doctests for the flat vector-query path",
    "l2_doctest": {
        "source": "python/python/lancedb/table.py:3587-3603",
        "vectors": [[0.1, 2.3, 4.5], [0.5, 3.4, 1.3], [0.3, 6.2, 2.6]], "query": [0.4, 1.4, 2.4],
        "captions_in_order": ["foo", "bar", "test"], "distances_6dp": ["5.220000", "5.309999", "23.089996"],
        "note": "the doctest prints rows 1 and 2 after a filter; 5.309999 is the oracle's value for the row it omits"},
    "cosine_doctest": {
        "source": "python/python/lancedb/query.py:1555-1571",
        "vectors": [[1.1, 1.2], [0.5, 1.3], [0.4, 0.4], [0.4, 0.4]], "b": [2, 4, 6, 10], "query": [0.4, 0.4],
        "b_in_order": [6, 10, 2], "distances_6dp": ["0.000000", "0.000000", "0.000944"]},
    "tie_break": {"source": "python/python/lancedb/query.py:1364-1370", "rule": "(_distance ASC, _rowid ASC)"},
    "exact_match": {"source": "python/python/tests/test_db.py:198-199", "distance": 0.0},
    "cosine_formula": {"source": "python/python/tests/test_query.py:993-1014", "tolerance": 1e-6},
}


def lance_fixture():
    from lancedb_b200 import lance_index
    from lancedb_b200.index import IvfPqIndexData
    z = np.load(os.path.join(HERE, "ivfpq_small.npz"))
    ix = IvfPqIndexData(32, 8, 4, "cosine", z["cosine_centroids"], z["cosine_codebook"], z["cosine_part_offsets"],
                        z["cosine_codes_t"], z["cosine_row_ids"], None)
    lance_index.write_ivf_pq_index(os.path.join(HERE, "lance_ivfpq_small"), ix, transposed=True, page_rows=256)


def main():
    if "--all" not in sys.argv:
        lance_fixture()
        print("wrote lance_ivfpq_small/ (pass --all to regenerate the oracle fixtures too)")
        return
    import oracle
    from tests.util import queries, random_index
    with open(os.path.join(HERE, "reference_pins.json"), "w") as f:
        json.dump(PINS, f, indent=1)
        f.write("\n")
    out = {}
    for metric in ("l2", "cosine", "dot"):
        rng = np.random.default_rng({"l2": 101, "cosine": 102, "dot": 103}[metric])
        ix = random_index(rng, dim=32, nlist=8, m=4, metric=metric, n=600, with_vectors=True)
        q = queries(rng, 6, 32)
        orc = oracle.OracleIndex.from_data(ix)
        allowed = np.sort(rng.choice(600, 90, replace=False)).astype(np.uint64)
        bm = oracle.allow_bitmap(allowed, 600)
        out.update({f"{metric}_centroids": ix.centroids, f"{metric}_codebook": ix.codebook,
                    f"{metric}_part_offsets": ix.part_offsets, f"{metric}_codes_t": ix.codes_t,
                    f"{metric}_row_ids": ix.row_ids, f"{metric}_vectors": ix.vectors, f"{metric}_queries": q,
                    f"{metric}_allow": bm})
        cases = {"plain": dict(k=7, nprobes=3), "range": dict(k=7, nprobes=3, lower=2.0, upper=30.0),
                 "refine": dict(k=5, nprobes=3, refine_factor=3), "prefilter": dict(k=7, nprobes=4, allow=bm, allow_bits=600)}
        for name, kw in cases.items():
            ids, dist, cnt = orc.search(q, **kw)
            out[f"{metric}_{name}_ids"], out[f"{metric}_{name}_dist"], out[f"{metric}_{name}_cnt"] = ids, dist, cnt
        fi, fd, fc = oracle.flat_search(ix.vectors, q, k=7, metric=metric, row_ids=ix.row_ids)
        out[f"{metric}_flat_ids"], out[f"{metric}_flat_dist"], out[f"{metric}_flat_cnt"] = fi, fd, fc
    np.savez_compressed(os.path.join(HERE, "ivfpq_small.npz"), **out)
    lance_fixture()
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
