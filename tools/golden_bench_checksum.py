"""Whole-batch checksums of the bench workload (50 000 config-2 loci, the bench seed) from the compiled reference and from
the oracle port, 8 host threads, ~10 CPU-minutes: contig count, DP cells, alignment score sum, trim/QC/selection checksum.
bench.py compares every run against tests/golden/bench_config2_checksum.json (run where /root/reference exists)."""
import ctypes as C, os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manta_b200 import _cabi as cb, synth  # noqa: E402
loci, seed = 50000, 20260923
batch = synth.make_candidate_batch(loci, seed=seed)
opts = cb.default_candidate_opts()
cuts = cb.default_ref_cuts(batch.nLoci)
res = {}
for name, path, prefix in [("reference", os.path.join(ROOT, "oracle", "_ref", "libmanta_ref.so"), "ref_"), ("oracle", os.path.join(ROOT, "oracle", "_build", "libmanta_oracle.so"), "orc_")]:
    b = cb.Binding(C.CDLL(path), prefix)
    t0 = time.time()
    got = b.candidates_mt(opts, batch, batch.nLoci, 8, cuts, True)
    res[name] = [int(x) for x in got]
    print(name, got, time.time() - t0, flush=True)
    json.dump(dict(loci=loci, seed=seed, results=res), open(os.path.join(ROOT, "tests", "golden", "bench_config2_checksum.json"), "w"), indent=1)
