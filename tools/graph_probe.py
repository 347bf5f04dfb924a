"""Does CUDA-graph replay of one UNet step's launch program beat stream launches?  (inter-kernel gaps at small batch)
usage: python tools/graph_probe.py [batch ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vq_voice_swap_b200 import engine, lib as L

dev = torch.device("cuda:0")
model = bench.build_model(dev, 64)
for batch in [int(a) for a in sys.argv[1:]] or [1, 4, 64]:
    x = torch.randn(batch, 1, 64000, device=dev)
    ts = torch.full((batch,), 0.5, device=dev)
    plan = engine._predictor_plan(model.predictor, x, None)
    engine.stage_predictor_inputs(model.predictor, plan, x, ts, None, None)
    co = plan.slots["conv_out"]
    co.mode, co.out = L.OUT_EPS, plan.eps.data_ptr()
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run()
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / reps
    ref = plan.eps.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        plan.run()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan.run()
    g.replay(); torch.cuda.synchronize()
    same = torch.equal(ref, plan.eps)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / reps
    print("batch %d: stream launches %.3f ms per UNet step, graph replay %.3f ms (%.1f %%), identical output: %s" % (batch, eager, graph, 100 * (graph / eager - 1), same), flush=True)
    del plan, x
    model.predictor._plans.clear()
    torch.cuda.empty_cache()
