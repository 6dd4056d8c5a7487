"""2-GPU data-parallel correctness (run by tests/test_gpu_multi.py through torch.distributed.run): two ranks with half the
batch each must reproduce the gradient of the GLOBAL mean loss -- BatchNorm statistics per replica, as the towers of
keras.utils.multi_gpu_model (learn_image_embeddings.py:133,148) -- and end the step with bit-identical weights.
The oracle is therefore run per shard (per-shard BatchNorm) and its gradients summed.  argv[1]: 'native' (the library's
NCCL communicator, bucketed all-reduces inside the captured step graph) or 'torch' (torch.distributed all_reduce)."""
import os, subprocess, sys
if 'RANK' not in os.environ:
    sys.exit(subprocess.call([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                              '--master-addr', '127.0.0.1', '--master-port', '29513', __file__] + sys.argv[1:]))
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import models as omodels, train as otrain
from semantic_embeddings_b200 import _lib, utils
from semantic_embeddings_b200.engine import Engine
from semantic_embeddings_b200.parallel import init_process_group
comm = sys.argv[1] if len(sys.argv) > 1 else 'native'
local = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(local)
rank, world = init_process_group(device=torch.device('cuda', local))
emb = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'class_matrices.npz'))['cifar100_embedding']
B, arch = 8, 'resnet-32'
emb = np.eye(64) if arch == 'resnet-32' else emb
om = omodels.build_network(emb.shape[1], arch, input_channels=3, seed=3)
omodels.randomize(om, seed=4)
for v in om.params.values(): v.copy_(v.float().double())
eng = Engine(utils.build_network(emb.shape[1], arch, input_channels=3), B, emb, device='cuda:%d' % local, world_size=world,
             mode=_lib.SE_MODE_TF32X3, comm=comm, use_cuda_graph=True)
assert eng.comm_native == (comm == 'native')
eng.set_weights({k: v.numpy().astype(np.float32) for k, v in om.params.items()})
g = torch.Generator().manual_seed(5)
C = emb.shape[0]
x = torch.randn(world * B, 32, 32, 3, generator=g).float(); y = torch.randint(0, C, (world * B,), generator=g)
eng.train_step(x[rank * B:(rank + 1) * B], y[rank * B:(rank + 1) * B], lr=0.05)
got = eng.get_grads()
# oracle: sum over shards of grad( (1/global_B) * sum_local loss ) with per-shard BN, + L2 term once
emb_t = torch.as_tensor(emb.astype(np.float32)).double()
tot = None
for r in range(world):
    leaf = {n: om.params[n].detach().clone().requires_grad_(n in om.trainable) for n in om.params}
    obj = otrain.train_objective(om, x[r * B:(r + 1) * B].double(), y[r * B:(r + 1) * B], emb_t, params=leaf)
    loss = obj['per_sample'].sum() / (world * B)
    gl = torch.autograd.grad(loss, [leaf[n] for n in om.trainable], allow_unused=True)
    gl = [t if t is not None else torch.zeros_like(leaf[n]) for t, n in zip(gl, om.trainable)]
    tot = gl if tot is None else [a + b for a, b in zip(tot, gl)]
num = den = 0.0
for n, gref in zip(om.trainable, tot):
    gref = gref + 2 * om.l2.get(n, 0.0) * om.params[n]
    a = got[n].astype(np.float64); b = gref.numpy()
    num += ((a - b) ** 2).sum(); den += (b ** 2).sum()
err = float(np.sqrt(num / den))
# a second step exercises the replay of the captured graph (all-reduces included)
eng.train_step(x[rank * B:(rank + 1) * B], y[rank * B:(rank + 1) * B], lr=0.05)
w0 = eng.P.clone(); dist.broadcast(w0, 0)
same = bool(torch.equal(w0, eng.P))
print('rank', rank, 'comm', comm, 'grad rel err vs per-shard-BN oracle %.3e' % err, 'weights identical across ranks:', same, flush=True)
assert err < 5e-3 and same
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print('DP_CHECK_OK', comm, flush=True)
