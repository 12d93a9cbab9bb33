"""A sharded handle with world = 1 on one GPU (gbp_cuda_init_shard -> gbp_shard_build_view -> build) against the plain
handle: every tensor bit-identical after a few sweeps."""
import os, socket, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
with socket.socket() as _s:
    _s.bind(("127.0.0.1", 0))
    _port = _s.getsockname()[1]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
import torch, torch.distributed as dist
import common
from gbp_poplar_b200 import GBPEngine, default_opts
from gbp_poplar_b200.engine import TENSOR_NAMES
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
st = common.make_setup("fr1xyz")
a = GBPEngine(st.problem, default_opts())
b = GBPEngine.sharded(st.problem, default_opts())
for e in (a, b):
    common.run_ba(e, 14)
    e.iterate(9)
bad = [t for t in TENSOR_NAMES if a.get_tensor(t).tobytes() != b.get_tensor(t).tobytes()]
print("world-1 sharded vs plain:", "IDENTICAL" if not bad else ("DIFFERENT " + str(bad)), flush=True)
b.close(); a.close()
dist.destroy_process_group()
