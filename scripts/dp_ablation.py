"""Where the data-parallel train step's time goes (B = 256 per GPU, D = 1728): run under torchrun.
  BN exchange: peer (NVLink peer-memory kernel) | nccl | none (local BatchNorm statistics: not the reference semantics, timing only)
  gradients  : bucket (two all-reduces inside the step) | flat (one after the step) | none (timing only)
torchrun --nproc-per-node 2 scripts/dp_ablation.py"""
import argparse, os, sys, time, types
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200 import train as T
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.optim import Adam
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = torch.device(f"cuda:{local}")
dist.all_reduce(torch.zeros(1, device=dev))
D, B = 1728, int(os.environ.get("B", 256))
x, _ = synth_windows(B, D, 1234 + rank, anomaly_rate=0.0)
x = x.to(dev)


def run(bn, grads, steps=40):
    cfg = argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=local, precision="f16x3")
    model = get_model(cfg)
    model.load_state_dict(synth_state_dict(D, 100, 5, 0))
    opt = Adam(model.parameters(), lr=1e-3)
    if bn != "none":
        T.set_data_parallel(model, peer=(bn == "peer"), overlap_grads=(grads == "bucket"))
    st = T.train_state(model)

    def step():
        model.train(); opt.zero_grad()
        loss = model.get_loss_value(x, None)
        loss.backward()
        if grads == "flat":
            if bn == "none":
                dist.all_reduce(st.flat_grad)
            else:
                T.allreduce_gradients(model)
        opt.step()
        return float(loss.detach())
    for _ in range(5):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world} B {B}: BN {bn:5s} grads {grads:6s}: {float(t):.3f} ms/step", flush=True)


cases = (("none", "none"), ("none", "flat"), ("nccl", "none"), ("peer", "none"), ("nccl", "flat"), ("peer", "flat"), ("nccl", "bucket"), ("peer", "bucket"))
if os.environ.get("CASES"):      # e.g. CASES=peer:flat,nccl:flat
    cases = tuple(tuple(c.split(":")) for c in os.environ["CASES"].split(","))
for bn, grads in cases:
    if grads == "bucket" and bn == "none":
        continue
    run(bn, grads)
dist.destroy_process_group()
