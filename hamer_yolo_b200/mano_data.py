"""MANO model data: loader for the official ``MANO_RIGHT.pkl`` and a seeded synthetic stand-in.

The reference gets these arrays through ``smplx.MANOLayer.__init__`` (third-party, called from
hamer/hamer/models/mano_wrapper.py:12-30 with cfg.MANO.MODEL_PATH).  smplx and chumpy are not
dependencies here: the pickle is read with a stub unpickler that maps every ``chumpy`` object onto a
plain attribute bag and resolves the one ``Select`` node (``shapedirs``) by hand.

Array conventions (what ``hm_set_mano`` in include/hamer_b200.h expects):
  v_template (778,3) f32     shapedirs (778,3,10) f32     posedirs (135,2334) f32
  J_regressor (16,778) f32   lbs_weights (778,16) f32     parents (16,) i32 (root = -1)
  faces (1538,3) i32         tip_ids (5,) i32             joint_map (21,) i32
"""
import pickle

import numpy as np

NUM_VERTS = 778
NUM_JOINTS = 16
NUM_FACES = 1538
# smplx.vertex_ids.vertex_ids['mano']: thumb, index, middle, ring, pinky (third-party constant)
TIP_IDS = np.array([744, 320, 443, 554, 671], dtype=np.int32)
# mano_wrapper.py:24
JOINT_MAP = np.array([0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20],
                     dtype=np.int32)
PARENTS = np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14], dtype=np.int32)


class _Bag:
    """Stand-in for any chumpy class found in the pickle."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.state = state


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith('chumpy'):
            return _Bag
        return super().find_class(module, name)


def _to_array(v):
    """Resolve a (possibly chumpy) node to a numpy array."""
    if isinstance(v, np.ndarray):
        return v
    if hasattr(v, 'toarray'):
        return np.asarray(v.toarray())
    if isinstance(v, _Bag):
        d = v.__dict__
        if 'x' in d:                      # chumpy.Ch
            return np.asarray(_to_array(d['x']))
        if 'a' in d and 'idxs' in d:      # chumpy.reordering.Select(a, idxs) with preferred_shape
            a = _to_array(d['a']).ravel()
            out = a[np.asarray(d['idxs'])]
            shape = d.get('preferred_shape')
            return out.reshape(shape) if shape is not None else out
    return np.asarray(v)


def load_mano_pkl(path, num_betas=10):
    """Read MANO_RIGHT.pkl / MANO_LEFT.pkl and return the arrays in the layout smplx builds
    (shapedirs[..., :num_betas]; posedirs reshaped to (135, V*3); parents[0] = -1)."""
    with open(path, 'rb') as f:
        d = _Unpickler(f, encoding='latin1').load()
    v_template = _to_array(d['v_template']).astype(np.float32)
    shapedirs = _to_array(d['shapedirs']).astype(np.float32)[:, :, :num_betas]
    posedirs = _to_array(d['posedirs']).astype(np.float32)
    posedirs = posedirs.reshape(-1, posedirs.shape[-1]).T.copy()        # (135, 2334)
    J_regressor = _to_array(d['J_regressor']).astype(np.float32)
    weights = _to_array(d['weights']).astype(np.float32)
    kintree = np.asarray(d['kintree_table'])
    parents = kintree[0].astype(np.int64).copy()
    parents[0] = -1
    return {
        'v_template': np.ascontiguousarray(v_template),
        'shapedirs': np.ascontiguousarray(shapedirs),
        'posedirs': np.ascontiguousarray(posedirs),
        'J_regressor': np.ascontiguousarray(J_regressor),
        'lbs_weights': np.ascontiguousarray(weights),
        'parents': parents.astype(np.int32),
        'faces': np.asarray(d['f']).astype(np.int32),
        'tip_ids': TIP_IDS.copy(),
        'joint_map': JOINT_MAP.copy(),
    }


def synthetic_mano(seed=0):
    """A seeded hand-like model with MANO's shapes, kinematic tree, sparsity and magnitudes
    (|shapedirs| ~ 6e-3, |posedirs| ~ 6e-3, <= 6 skinning weights per vertex, ~1.9k J_regressor
    non-zeros).  Used wherever the licensed MANO pickle is unavailable (tests on the GPU box,
    bench.py); the kernels and the oracle are handed the same arrays."""
    rng = np.random.default_rng(seed)
    # skeleton in metres: wrist + 5 fingers x 3 joints, MANO order index, middle, pinky, ring, thumb
    J = np.zeros((NUM_JOINTS, 3), np.float64)
    spread = np.array([-0.035, -0.012, 0.030, 0.010, -0.045])
    length = np.array([0.032, 0.034, 0.024, 0.030, 0.030])
    base = np.array([0.090, 0.095, 0.080, 0.088, 0.035])
    for f in range(5):
        for k in range(3):
            J[1 + 3 * f + k] = [base[f] + k * length[f], spread[f] * (1 + 0.15 * k), 0.004 * (f - 2)]
    # vertices: palm cloud around the wrist..knuckles box plus points scattered along each bone
    bone_child = np.arange(1, NUM_JOINTS)
    verts = []
    n_palm = 200
    palm = np.stack([rng.uniform(-0.01, 0.09, n_palm), rng.uniform(-0.045, 0.04, n_palm),
                     rng.normal(0, 0.008, n_palm)], axis=1)
    verts.append(palm)
    per_bone = (NUM_VERTS - n_palm) // len(bone_child)
    extra = (NUM_VERTS - n_palm) - per_bone * len(bone_child)
    for bi, c in enumerate(bone_child):
        n = per_bone + (1 if bi < extra else 0)
        t = rng.uniform(0.0, 1.2, n)[:, None]
        seg_end = J[c] + (J[c] - J[PARENTS[c]]) * (0.9 if c % 3 == 0 else 0.0)
        p = J[c] * (1 - t) + (seg_end if c % 3 == 0 else J[min(c + 1, NUM_JOINTS - 1)]) * t
        verts.append(p + rng.normal(0, 0.006, (n, 3)))
    v_template = np.concatenate(verts, axis=0)[:NUM_VERTS]
    assert v_template.shape == (NUM_VERTS, 3)
    # skinning weights: softmax of -d^2 over joints, top-6 kept, renormalised
    d2 = ((v_template[:, None, :] - J[None]) ** 2).sum(-1)
    w = np.exp(-d2 / (2 * 0.012 ** 2))
    order = np.argsort(-w, axis=1)
    mask = np.zeros_like(w, dtype=bool)
    np.put_along_axis(mask, order[:, :6], True, axis=1)
    w = np.where(mask & (w > 0.02 * w.max(axis=1, keepdims=True)), w, 0.0)   # mean ~2.6 non-zeros like MANO
    w /= w.sum(axis=1, keepdims=True)
    # fingertip vertices are rigidly attached to the distal joints (as in MANO)
    for tip, joint in zip(TIP_IDS, (15, 3, 6, 12, 9)):
        w[tip] = 0.0
        w[tip, joint] = 1.0
        v_template[tip] = J[joint] + (J[joint] - J[PARENTS[joint]]) * 0.9
    # joint regressor: convex combination of the ~120 nearest vertices
    Jr = np.zeros((NUM_JOINTS, NUM_VERTS))
    for j in range(NUM_JOINTS):
        dj = ((v_template - J[j]) ** 2).sum(-1)
        idx = np.argsort(dj)[:118]
        a = rng.uniform(0.2, 1.0, idx.size)
        Jr[j, idx] = a / a.sum()
    shapedirs = rng.normal(0, 2.0e-3, (NUM_VERTS, 3, 10))
    posedirs = rng.normal(0, 1.5e-3, (135, NUM_VERTS * 3))
    faces = rng.integers(0, NUM_VERTS, (NUM_FACES, 3))
    return {
        'v_template': v_template.astype(np.float32),
        'shapedirs': shapedirs.astype(np.float32),
        'posedirs': posedirs.astype(np.float32),
        'J_regressor': Jr.astype(np.float32),
        'lbs_weights': w.astype(np.float32),
        'parents': PARENTS.copy(),
        'faces': faces.astype(np.int32),
        'tip_ids': TIP_IDS.copy(),
        'joint_map': JOINT_MAP.copy(),
    }
