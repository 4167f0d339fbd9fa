#!/bin/sh
python - <<'PY'
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import zpaq_v_b200 as z
import test_gpu_determinism as t
ctx = z.Context()
pr = cProfile.Profile(); pr.enable()
t0=time.time()
t.test_same_bytes_under_every_packing.__wrapped__(ctx, 2) if hasattr(t.test_same_bytes_under_every_packing,'__wrapped__') else t.test_same_bytes_under_every_packing(ctx, 2)
print("total", time.time()-t0)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
PY
