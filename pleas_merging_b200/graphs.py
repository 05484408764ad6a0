"""CUDA-graph replay of a per-batch pipeline.

A calibration step is ~1 700 kernel launches for a ResNet-50 pair (two forwards plus four of
this library's kernels per tap / six per layer); a Python launch loop cannot issue them as fast
as a B200 retires them.  ``GraphedStep`` runs the first batch of each input shape eagerly (it
binds staging memory and warms cuDNN), captures the second into a CUDA graph, and replays the
graph for every later batch through a static input buffer.  Capture is an optimisation only:
the same kernels run either way, and a failed capture falls back to eager launches.
"""
import warnings

import torch


class GraphedStep:
    def __init__(self, fn, after_first=None, use_cuda_graph=True):
        """fn(x) enqueues one batch's work on the current stream; after_first() runs once after
        the first eager batch of a shape (e.g. to rebind staging that grew during it)."""
        self.fn, self.after_first, self.use_cuda_graph = fn, after_first, use_cuda_graph
        self.graphs = {}

    def clear(self):
        self.graphs.clear()

    def __call__(self, x):
        if not self.use_cuda_graph:
            return self.fn(x)
        key = (tuple(x.shape), x.dtype)
        entry = self.graphs.get(key)
        if entry is None:
            self.fn(x)
            if self.after_first is not None:
                self.after_first()
            self.graphs[key] = "warm"
            return
        if entry == "warm":
            static_x = x.clone()
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            try:
                # capture_begin / capture_end directly instead of the torch.cuda.graph context manager:
                # that one runs gc.collect() and torch.cuda.empty_cache() first, i.e. it cudaFree()s the
                # multi-GB activation blocks the eager batch just cached and the capture then has to
                # cudaMalloc them again (0.1-0.3 s per capture on a ResNet-50 pair; 180 GB of HBM do not
                # need the memory back)
                side = torch.cuda.Stream(x.device)
                side.wait_stream(torch.cuda.current_stream(x.device))
                with torch.cuda.stream(side):
                    graph.capture_begin()
                    try:
                        self.fn(static_x)
                    finally:
                        graph.capture_end()
                torch.cuda.current_stream(x.device).wait_stream(side)
            except Exception as e:
                warnings.warn(f"CUDA-graph capture failed ({e}); running eagerly")
                self.use_cuda_graph = False
                torch.cuda.synchronize()
                return self.fn(x)
            self.graphs[key] = (graph, static_x)
            graph.replay()
            return
        graph, static_x = entry
        static_x.copy_(x, non_blocking=True)
        graph.replay()
