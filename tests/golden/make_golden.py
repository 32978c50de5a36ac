"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = the reference's search_worker.cc,
linear_search.cc, build_hash_tables.cc compiled unmodified; see oracle/ref_driver.cc).  Needs /root/reference
(to build oracle/_ref); the fixtures are committed so that the GPU box, which has no reference tree, can check
against them.  Inputs are regenerated from the seeds by the shared synthetic generator; a CRC of the code array
is stored to pin it.

    python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference as F, restatement as R  # noqa: E402
import oracle  # noqa: E402

CASES = [
    # name, n, bits, m, k, nq, approximate
    ("mih64_m4_k10", 20000, 64, 4, 10, 8, False),
    ("mih64_m4_k100", 50000, 64, 4, 100, 4, False),
    ("mih128_m8_k10", 4000, 128, 8, 10, 3, False),
    ("mih64_m4_k10_approx", 8000, 64, 4, 10, 6, True),
    ("mih64_m4_k100_small", 60, 64, 4, 100, 3, False),
]
DB_SEED, Q_SEED = 12345, 67890


def main():
    oracle.build()
    for name, n, bits, m, k, nq, approx in CASES:
        nbytes = bits // 8
        codes = R.synth_codes(DB_SEED, 0, n, nbytes)
        queries = R.synth_codes(Q_SEED, 0, nq, nbytes)
        store = F.RefStore(codes, m)                      # reference build-tables main(), per rank
        ids, dists, counts, radius, sub_reads = store.mih_search(queries, k, approximate=approx)
        store.put_main_table()
        lids, ldists, lcounts = store.linear_search(queries, k)
        # a few buckets as stored by the reference build
        probe_keys = np.array([R.binary_to_int(codes[i, : nbytes // m]) for i in (0, n // 2, n - 1)], dtype=np.uint32)
        bucket_ids = [store.bucket(0, int(key))[1] for key in probe_keys]
        store.close()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            n=n, bits=bits, m=m, k=k, nq=nq, approximate=approx, db_seed=DB_SEED, q_seed=Q_SEED,
            codes_crc=np.uint32(zlib.crc32(codes.tobytes())), queries=queries,
            ref_mih_ids=ids, ref_mih_dists=dists, ref_mih_counts=counts, ref_mih_radius=radius, ref_mih_sub_reads=sub_reads,
            ref_lin_ids=lids, ref_lin_dists=ldists, ref_lin_counts=lcounts,
            probe_keys=probe_keys, bucket0=bucket_ids[0], bucket1=bucket_ids[1], bucket2=bucket_ids[2],
        )
        print("wrote", name, "radius", sorted(set(radius.tolist())))


if __name__ == "__main__":
    main()
