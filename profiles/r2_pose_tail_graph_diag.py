"""Which call of the PoseTail train step invalidates a stream capture?  Captures eval forward, train forward, train
forward + backward separately and prints the FIRST exception of each (torch.cuda.graph.__exit__ masks it otherwise)."""
import sys, traceback, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as ge
pkg = ge.load_package(); dev = torch.device('cuda:0')
tail = pkg.PoseTail(1024, 1024).to(dev)
B = 138
sf = torch.randn(B, 1024, device=dev, requires_grad=True); img = torch.randn(B, 1024, device=dev, requires_grad=True)


def attempt(name, fn, warm=3):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            tail.zero_grad(set_to_none=True); sf.grad = None; img.grad = None
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    tail.zero_grad(set_to_none=True); sf.grad = None; img.grad = None
    g = torch.cuda.CUDAGraph()
    first = []
    try:
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            try:
                fn()
            except BaseException as e:      # the first failure, before capture_end can mask it
                first.append(traceback.format_exc())
                raise
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: captured OK, {e0.elapsed_time(e1) * 1000 / 50:.1f} us per replay", flush=True)
    except BaseException as e:
        print(f"{name}: FAILED: {type(e).__name__}: {str(e)[:200]}", flush=True)
        for t in first:
            print("  first exception inside the capture:\n" + t[-1500:], flush=True)
    torch.cuda.synchronize()


def fwd_eval():
    with torch.no_grad():
        return tail(sf, img)


def fwd_train():
    return tail(sf, img)


def fwd_bwd():
    outs, x, p = tail(sf, img)
    (sum(o.sum() for o in outs) + x.sum() + p.sum()).backward()


tail.eval(); attempt("eval forward", fwd_eval)
tail.train(); attempt("train forward", fwd_train)
attempt("train forward + backward", fwd_bwd)
