"""Where does the end-to-end step time go?  Device-resident eager / graph replay / per-step sync / host copies."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import bench


def timed(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def main():
    m = bench.build_model().cuda()
    eng = m._engine()
    B = 32
    host_in = torch.randn(B, 1, 8000).pin_memory()
    host_out = torch.empty(B, 1, 32000).pin_memory()
    x = host_in.cuda()
    m.use_cuda_graph(False)
    for _ in range(3):
        m(x)
    print("eager, no sync        : %.2f ms (wall %.2f)" % timed(lambda: m(x)))

    def sync_step():
        m(x)
        torch.cuda.current_stream().synchronize()
    print("eager, sync per step  : %.2f ms (wall %.2f)" % timed(sync_step))
    m.use_cuda_graph(True)
    for _ in range(3):
        m(x)
    print("graph, no sync        : %.2f ms (wall %.2f)" % timed(lambda: m(x)))
    print("graph, sync per step  : %.2f ms (wall %.2f)" % timed(sync_step))

    def e2e():
        xin = host_in.to("cuda", non_blocking=True)
        out = m(xin)
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    print("graph, e2e            : %.2f ms (wall %.2f)" % timed(e2e))
    key, (graph, si, so) = next(iter(eng._graphs.items()))
    print("raw graph.replay      : %.2f ms (wall %.2f)" % timed(graph.replay))


if __name__ == "__main__":
    main()
