"""bf16-bank variants of the CRD scoring kernel at the headline config (kernel ms): round 1 measured {1: 0.506, 2 (now the
default): 0.367, 3: 0.906}; four deeper / wider shapes tried and dropped (0.54-0.65 ms, register spills)."""
import sys, torch, json
sys.path.insert(0, '.')
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
dev = torch.device('cuda:0')
res = {}
for v in (0, 1, 3):
    try:
        r = bench.time_crd_resident(pkg, torch, dev, bench.HEADLINE, 30, 3, variant=v, bank_dtype=torch.bfloat16)
        res[v] = round(r["kernel_ms_avg"], 4)
    except Exception as e:
        res[v] = str(e)[:80]
print(json.dumps(res))
