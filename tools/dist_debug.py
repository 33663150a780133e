"""2-rank NCCL smoke of the bucketed all-reduce, eager and captured inside a CUDA graph (launch with torchrun)."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "50")), exit=True)
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
def log(*a):
    print(f"[r{rank} {time.time() % 1000:.2f}]", *a, file=sys.stderr, flush=True)
from _models import build_dsdgp
model, X, Y = build_dsdgp(L=3, M=40, D=4, N=601, S=3, minibatch=151)
model.compile(f"cuda:{local}")
model.distribute()
mode = os.environ.get("MODE", "graph")
model.use_cuda_graph = mode == "graph"
log("eager/first call, mode", mode)
for i in range(3):
    v, g = model.elbo_and_grad()
    torch.cuda.synchronize()
    log("step", i, float(v))
dist.barrier()
model.release_graphs()
log("done")
dist.destroy_process_group()
