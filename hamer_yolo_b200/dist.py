"""Data-parallel sharding of hand crops across the GPUs of one box (one process per GPU).

The reference has no multi-GPU path for inference (single process, batch 1 per hand,
hamer/infer.py:1268-1274); crops are independent, so the path shards with NO data-path exchange
except one collective that collects the per-hand MANO outputs (north_star): an all-gather of a
packed record per hand over NCCL / NVLink.

Record layout (float32, RECORD_FLOATS = 2599 values = 10,396 bytes per hand):
  verts 778*3 | kp3d 21*3 | kp2d 21*2 | cam 3 | cam_t 3 | rotmats 16*9 | betas 10
"""
import ctypes as C

import torch
import torch.distributed as dist

FIELDS = (('pred_vertices', 778 * 3), ('pred_keypoints_3d', 63), ('pred_keypoints_2d', 42),
          ('pred_cam', 3), ('pred_cam_t', 3), ('rotmats', 144), ('betas', 10))
RECORD_FLOATS = sum(n for _, n in FIELDS)


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced partition: rank r owns [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_outputs(output):
    """HAMER.forward output dict -> [B, RECORD_FLOATS] float32 (same device)."""
    mp = output['pred_mano_params']
    B = output['pred_cam'].shape[0]
    rot = torch.cat([mp['global_orient'].reshape(B, 9), mp['hand_pose'].reshape(B, 135)], dim=1)
    parts = [output['pred_vertices'].reshape(B, -1), output['pred_keypoints_3d'].reshape(B, -1),
             output['pred_keypoints_2d'].reshape(B, -1), output['pred_cam'], output['pred_cam_t'],
             rot, mp['betas']]
    rec = torch.cat(parts, dim=1).contiguous()
    assert rec.shape[1] == RECORD_FLOATS
    return rec


def unpack_outputs(rec):
    """Inverse of pack_outputs (plus the split of rotmats into global_orient / hand_pose)."""
    B = rec.shape[0]
    out, off = {}, 0
    for name, n in FIELDS:
        out[name] = rec[:, off:off + n]
        off += n
    rot = out.pop('rotmats').reshape(B, 16, 3, 3)
    betas = out.pop('betas')
    return {
        'pred_vertices': out['pred_vertices'].reshape(B, 778, 3),
        'pred_keypoints_3d': out['pred_keypoints_3d'].reshape(B, 21, 3),
        'pred_keypoints_2d': out['pred_keypoints_2d'].reshape(B, 21, 2),
        'pred_cam': out['pred_cam'], 'pred_cam_t': out['pred_cam_t'],
        'pred_mano_params': {'global_orient': rot[:, :1], 'hand_pose': rot[:, 1:], 'betas': betas},
    }


def init_comm(engine, group=None):
    """Give ``engine`` (one hm_ctx = one rank = one GPU) its own NCCL communicator through the C ABI: rank 0 creates the
    id (hm_comm_unique_id), torch.distributed only ferries the 128 bytes, every rank joins with hm_comm_init.  After this,
    all_gather_records(..., engine=engine) is ncclAllGather issued by libhamer_b200.so (hm_allgather)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ident = [None]
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        rc = engine.lib.hm_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f'hm_comm_unique_id failed ({rc}): {engine.lib.hm_last_error(None).decode()}')
        ident = [bytes(buf)]
    dist.broadcast_object_list(ident, src=0, group=group)
    raw = (C.c_uint8 * 128).from_buffer_copy(ident[0])
    with torch.cuda.device(engine.device):
        engine.check(engine.lib.hm_comm_init(engine.h, raw, world, rank))
    engine.comm_ranks = world
    return engine


def all_gather_records(rec_local, counts, group=None, engine=None):
    """Gather every rank's [B_r, RECORD_FLOATS] block into [sum(B_r), RECORD_FLOATS] on all ranks.
    Equal counts are one ncclAllGather; ragged counts pad to the max.  With ``engine`` (after init_comm) the collective
    is issued by the library itself (hm_allgather) on the current stream, otherwise by torch.distributed."""
    world = getattr(engine, 'comm_ranks', 0) if engine is not None else 0
    if not world:
        engine = None
        world = dist.get_world_size(group)
    assert len(counts) == world
    bmax = max(counts)
    if rec_local.shape[0] != bmax:
        pad = torch.zeros(bmax - rec_local.shape[0], rec_local.shape[1], dtype=rec_local.dtype,
                          device=rec_local.device)
        rec_local = torch.cat([rec_local, pad], dim=0)
    out = torch.empty(world * bmax, rec_local.shape[1], dtype=rec_local.dtype, device=rec_local.device)
    if engine is not None:
        rec_local = rec_local.contiguous()
        assert rec_local.is_cuda and rec_local.dtype == torch.float32 and rec_local.shape[1] == RECORD_FLOATS
        with torch.cuda.device(engine.device):
            engine.check(engine.lib.hm_allgather(engine.h, None, C.c_void_p(rec_local.data_ptr()), C.c_void_p(out.data_ptr()),
                                                 bmax, engine.stream()))
    else:
        dist.all_gather_into_tensor(out, rec_local.contiguous(), group=group)
    if all(c == bmax for c in counts):
        return out
    return torch.cat([out[r * bmax:r * bmax + counts[r]] for r in range(world)], dim=0)


def forward_sharded(model, img_all, group=None):
    """Run ``model`` on this rank's contiguous shard of ``img_all`` ([N,3,256,256], same on every
    rank or only the local part valid) and all-gather the MANO outputs.  Returns the unpacked
    global dict (every rank holds all N hands)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = img_all.shape[0]
    lo, hi = shard_range(n, rank, world)
    counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    if hi > lo:
        out, _ = model({'img': img_all[lo:hi]})
        rec_local = pack_outputs(out)
    else:
        # fewer hands than ranks (2 hands of a video frame on 8 GPUs): this rank has nothing to run but must still take
        # part in the collective, with a zero-row record (all_gather_records pads it to the largest shard)
        rec_local = torch.zeros(0, RECORD_FLOATS, dtype=torch.float32, device=img_all.device)
    rec = all_gather_records(rec_local, counts, group, engine=getattr(model, 'engine', None))
    return unpack_outputs(rec)
