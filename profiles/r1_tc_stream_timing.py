"""Times the CRD step at the headline shape with bf16 banks: gather kernel vs the tensor-core streaming kernel
(ContrastMemory.streaming).  Usage: python profiles/r1_tc_stream_timing.py > gpurun_out/tc_stream_timing.json"""
import importlib, json, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d-augmented-contrastive-knowledge-distillation-for-image-based-object-pose-estimation_b200")
dev = torch.device("cuda:0")
N, K, B = 1_000_000, 65536, 46
out = {}
for dt in (torch.bfloat16,):
    mem = pkg.ContrastMemory(128, N, K, 0.07, 0.5, bank_dtype=dt).to(dev)
    g = torch.Generator().manual_seed(1)
    v1 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(dev)
    v2 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(dev)
    y = torch.randperm(N, generator=g)[:B].to(dev)
    idxs = [torch.randint(0, N, (B, K + 1), device=dev) for _ in range(8)]
    for c in idxs:
        c[:, 0] = y
    mem._freeze_z(v1, v2, idxs[0])
    hp = mem._host_params()
    modes = (("gather", False), ("tc_stream", True))
    if os.environ.get("TC_ONLY"):
        modes = (("tc_stream", True),)
    for name, streaming in modes:
        mem.streaming = streaming
        for i in range(5):
            r = mem._step(v1, v2, y, idxs[i % 8], hp.Z1, hp.Z2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = int(os.environ.get('TC_ITERS', '40'))
        e0.record()
        for i in range(n):
            r = mem._step(v1, v2, y, idxs[i % 8], hp.Z1, hp.Z2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f"{str(dt).split('.')[-1]}_{name}"] = {"ms_per_step": ms, "scores_per_sec": B * (K + 1) / ms * 1e3,
                                                   "loss": float(r[0][5].item())}
print(json.dumps(out))
