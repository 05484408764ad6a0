"""Experiment: two exact-fp32 ResNet-50 forwards (B=32) in one CUDA graph, same stream vs two streams."""
import time, torch, torchvision
torch.backends.cudnn.allow_tf32 = False; torch.backends.cudnn.benchmark = True
torch.manual_seed(0); m1 = torchvision.models.resnet50().eval().cuda()
torch.manual_seed(1); m2 = torchvision.models.resnet50().eval().cuda()
x = torch.randn(32, 3, 224, 224, device="cuda")
side = torch.cuda.Stream()
def seq(): m1(x); m2(x)
def par():
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(main)
    with torch.cuda.stream(side):
        side.wait_event(ev); m2(x); done = torch.cuda.Event(); done.record(side)
    m1(x); main.wait_event(done)
with torch.inference_mode():
    for name, fn in (("sequential", seq), ("two streams", par)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): fn()
        for _ in range(3): g.replay()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): g.replay()
        torch.cuda.synchronize()
        print(f"{name}: {(time.perf_counter()-t0)/20*1e3:.2f} ms per pair of forwards")
