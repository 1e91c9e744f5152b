"""N-rank test of the C-ABI collective (needs >= 2 GPUs on the box: `gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`).
Two processes, one GPU each, shard 7 hands (ragged: 4 + 3), run the model on their shard and gather the records with
hm_allgather (ncclAllGather issued by libhamer_b200.so over its own communicator)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('gloo')                      # torch.distributed only ferries the NCCL id (CPU)
from hamer_yolo_b200 import load_hamer, synthetic_mano
from hamer_yolo_b200 import dist as hd
from hamer_yolo_b200.weights import random_state_dict
dev = torch.device('cuda', rank)
model, _ = load_hamer(state_dict=random_state_dict(seed=0, device=dev), mano=synthetic_mano(0), device=dev, max_batch=8,
                      deterministic=True)
hd.init_comm(model.engine)
n = int(sys.argv[2])
img = torch.randn(n, 3, 256, 256, generator=torch.Generator().manual_seed(7)).to(dev)
got = hd.forward_sharded(model, img)                 # shard -> forward -> hm_allgather
torch.cuda.synchronize()
full, _ = model({'img': img})                        # every rank also runs the whole batch: the gathered result must equal it
assert got['pred_vertices'].shape == (n, 778, 3)
assert torch.equal(got['pred_vertices'], full['pred_vertices']), 'vertices differ'
assert torch.equal(got['pred_mano_params']['hand_pose'], full['pred_mano_params']['hand_pose'])
assert torch.equal(got['pred_keypoints_2d'], full['pred_keypoints_2d'])
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('n', [7, 1])
def test_sharded_forward_with_hm_allgather_two_ranks(tmp_path, n):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER)
    port = 29700 + (os.getpid() % 200) + n
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(n)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out.decode()[-3000:]
