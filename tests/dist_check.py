"""Two-rank check of the sharded paths (launched by tests/test_methods_gpu.py with torchrun when
>= 2 GPUs are visible): distributed activation matching / PLeaS accumulate the same cost
matrices, permutations and fitted weights as a single-GPU run over the same batches."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cudnn.allow_tf32 = False
    import pleas_merging_b200 as P
    from oracle import tinynet

    m1, m2 = tinynet.make_pair(12, 10)
    spec = P.get_permutation_spec(m1, ((1, 3, 16, 16),))
    m1, m2 = m1.cuda(), m2.cuda()
    loader = tinynet.make_loader(5, 4, 16)
    for accumulate in ("sum", "reference"):
        p1, c1 = P.activation_matching(spec, m1, m2, loader, 5, output_costs=True, accumulate=accumulate)
        p2, c2 = P.activation_matching(spec, m1, m2, loader, 5, output_costs=True, accumulate=accumulate,
                                       distributed=True)
        for k in spec:
            err = (c1[k] - c2[k]).abs().max() / c1[k].abs().max()
            assert err <= 1e-6, (accumulate, k, float(err))
            assert torch.equal(p1[k], p2[k]), (accumulate, k)
    tl = tinynet.make_loader(9, 4, 16, seed=5)
    outs = []
    for distributed in (False, True):
        m3 = P.partial_merge(spec, m1, m2, p1, c1, 0.5)
        P.train(tl, m1, m2, m3, spec, p1, c1, 0.5, False, 8, None, num_classes=10, model_type="rn18",
                distributed=distributed)
        outs.append({k: v.clone() for k, v in m3.state_dict().items()})
    for k in outs[0]:
        assert torch.allclose(outs[0][k], outs[1][k], rtol=1e-4, atol=1e-6), k
    dist.barrier()
    if rank == 0:
        print("dist_check ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
