import os, sys
sys.path.insert(0, os.getcwd())
from upcgen_b200 import capi
from upcgen_b200.config import named_config
P = named_config("cfg2")
g = capi.UpcGpu(P, 0)
g.prepare_tables()
for n in (1, 2, 4, 8, 16):
    for sh in sorted(set((0, n // 2, n - 1))):
        for rep in range(3):
            g.fill_lumi_shard(sh, n)
        st = g.fill_stats()
        print(f"shards {n:2d} shard {sh:2d}: integrals {st['qags_integrals']:8d} evals {st['qags_evals']/1e6:8.1f}M head {st['ms_qags_head']:7.3f} ms qags {st['ms_qags']:7.3f} flux {st['ms_flux']:7.3f} cells {st['ms_cells']:6.3f}  ns/eval(head) {st['ms_qags_head']*1e6/st['qags_head_evals']:.4f}")
g.close()
