"""Drop-in mirror of the reference's model API, backed by libhamer_b200.so.

Reference interfaces mirrored (names, argument meaning, returned dict layout, error behaviour):
  hamer/hamer/models/hamer.py:19-72    HAMER.__init__        -> HAMER(cfg, ...)
  hamer/hamer/models/hamer.py:99-156   HAMER.forward_step    -> HAMER.forward_step / forward
  hamer/hamer/models/hamer.py:269-277  HAMER.forward(batch)  -> (output, pred_mano_params)
  hamer/hamer/models/__init__.py:32-47 load_hamer(ckpt)      -> load_hamer(ckpt, ...)
PyTorch is used only to own device memory and hand pointers + the current stream to the C ABI.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .configs import default_config
from .mano_data import load_mano_pkl, synthetic_mano

IMG = 256


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _ManoHandle:
    """What callers read from ``model.mano`` (infer.py:148-152 uses ``.faces``)."""

    def __init__(self, data):
        self.data = data
        self.faces = data['faces']


BLUR_MAX_RADIUS = 32          # HM_BLUR_MAX_RADIUS (include/hamer_b200.h)


def gaussian_taps(sigma, truncate=4.0):
    """(radius, weights[33]) of skimage.filters.gaussian / scipy.ndimage.gaussian_filter1d for ``sigma``: radius =
    int(truncate * sigma + 0.5), weights[k] = exp(-0.5 k^2 / sigma^2) / sum - scipy's _gaussian_kernel1d, evaluated with numpy
    as scipy does so that the taps are the reference's to the last bit."""
    radius = int(truncate * float(sigma) + 0.5)
    if radius > BLUR_MAX_RADIUS:
        raise ValueError(f'anti-alias sigma {float(sigma):.2f} needs radius {radius} > {BLUR_MAX_RADIUS}')
    sigma2 = sigma * sigma            # in sigma's own type, like scipy: ViTDetDataset passes a numpy float32
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    w = np.zeros(BLUR_MAX_RADIUS + 1, np.float64)
    w[:radius + 1] = phi[radius:]
    return radius, w


class _DeviceFrame:
    """A host frame that hm_letterbox uploaded into the context's frame buffer: hand it to Engine.preprocess so the crops
    read the device copy (valid until the next host frame goes through letterbox / preprocess)."""

    def __init__(self, ptr, H, W, row_stride, host):
        self.ptr, self.shape, self.row_stride, self.host = ptr, (H, W, 3), row_stride, host


class Engine:
    """Owns one hm_ctx (one device, one stream user).  Thin: argument marshalling only."""

    MAX_GRAPHS = 8      # cached CUDA graphs (least recently used is dropped): each pins a [B,3,256,256] static input

    def __init__(self, device, max_batch=64, max_frame_bytes=1920 * 1080 * 3, gemm_cta_group=0,
                 attention_legacy=0, mano_legacy=0, use_graphs=True, deterministic=0, split_weights=0):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('hamer_b200 runs on a CUDA (sm_100a) device only; there is no CPU path')
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', self.index)
        self.max_batch = int(max_batch)
        self.max_frame_bytes = int(max_frame_bytes)
        cfg = _lib.HmConfig(max_batch=self.max_batch, max_frame_bytes=self.max_frame_bytes,
                            gemm_cta_group=int(gemm_cta_group), attention_legacy=int(attention_legacy),
                            mano_legacy=int(mano_legacy), deterministic=int(deterministic),
                            split_weights=int(split_weights))
        h = C.c_void_p()
        rc = self.lib.hm_create(C.byref(h), self.index, C.byref(cfg))
        if rc != 0:
            raise _lib.HmError(rc, self.lib.hm_last_error(None).decode())
        self.h = h
        # CUDA graphs of hm_forward, one per batch size (the path is ~310 launches; replaying a graph
        # removes the launch gaps - most of the latency at the 2-8 hands of a video frame)
        self.use_graphs = bool(use_graphs)
        self._graphs = {}
        self._profiling = False
        self.comm_ranks = 0        # > 0 once dist.init_comm gave this context its own NCCL communicator

    def close(self):
        self._graphs = {}
        if getattr(self, 'h', None):
            self.lib.hm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def check(self, rc):
        return _lib.check(self.h, rc)

    # -- weights / MANO --------------------------------------------------------------------------
    def _drop_graphs(self):
        """Captured graphs hold weight / MANO table pointers and tensor maps by value: any reload invalidates them."""
        if self._graphs:
            torch.cuda.synchronize(self.device)
            self._graphs = {}

    def set_weight(self, name, tensor):
        self._drop_graphs()
        t = tensor.detach().to('cpu', torch.float32).contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        self.check(self.lib.hm_set_weight(self.h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))

    def finalize_weights(self):
        self._drop_graphs()
        self.check(self.lib.hm_finalize_weights(self.h))

    def set_mano(self, m):
        self._drop_graphs()
        arr = lambda k, dt: np.ascontiguousarray(m[k], dtype=dt)
        keep = [arr('v_template', np.float32), arr('shapedirs', np.float32), arr('posedirs', np.float32),
                arr('J_regressor', np.float32), arr('lbs_weights', np.float32), arr('parents', np.int32),
                arr('tip_ids', np.int32), arr('joint_map', np.int32)]
        assert keep[0].shape == (778, 3) and keep[1].shape == (778, 3, 10) and keep[2].shape == (135, 2334)
        assert keep[3].shape == (16, 778) and keep[4].shape == (778, 16)
        self.check(self.lib.hm_set_mano(self.h, *[C.c_void_p(a.ctypes.data) for a in keep]))

    # -- stages ----------------------------------------------------------------------------------
    def _outputs(self, B, want_feats=False):
        dev = self.device
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        t = {'pred_cam': f(B, 3), 'global_orient': f(B, 1, 3, 3), 'hand_pose': f(B, 15, 3, 3),
             'betas': f(B, 10), 'pred_cam_t': f(B, 3), 'focal_length': f(B, 2),
             'pred_keypoints_3d': f(B, 21, 3), 'pred_vertices': f(B, 778, 3),
             'pred_keypoints_2d': f(B, 21, 2)}
        if want_feats:
            t['conditioning_feats'] = f(B, 192, 1280)
        return t

    @staticmethod
    def _struct(t, lo, hi):
        s = _lib.HmOutputs()
        for name, _ in _lib.HmOutputs._fields_:
            setattr(s, name, t[name][lo:hi].data_ptr() if name in t else None)
        return s

    _FIELDS = (('pred_cam', (3,)), ('global_orient', (1, 3, 3)), ('hand_pose', (15, 3, 3)), ('betas', (10,)),
               ('pred_cam_t', (3,)), ('focal_length', (2,)), ('pred_keypoints_3d', (21, 3)),
               ('pred_vertices', (778, 3)), ('pred_keypoints_2d', (21, 2)))

    @classmethod
    def _flat_floats(cls, B):
        """Size of the packed output buffer of a graph: every field starts on a 64-byte boundary."""
        return sum((B * int(np.prod(shp)) + 15) // 16 * 16 for _, shp in cls._FIELDS)

    def _views(self, flat, B):
        # every field starts 64-byte aligned whatever B is: hm_mano_lbs takes its fused kernel only for an 8-byte aligned
        # vertex pointer, and with tightly packed fields an odd batch put pred_vertices 4 bytes off (the three-kernel
        # path was taken for 1, 3, 5, ... hands: same tolerances, different bits, slower)
        t, off = {}, 0
        for name, shp in self._FIELDS:
            n = B * int(np.prod(shp))
            t[name] = flat[off:off + n].view(B, *shp)
            off += (n + 15) // 16 * 16
        return t

    def _forward_graph(self, img):
        B = img.shape[0]
        g = self._graphs.pop(B, None)
        if g is None:
            static_img = torch.empty(B, 3, IMG, IMG, device=self.device, dtype=torch.float32)
            static_flat = torch.empty(self._flat_floats(B), device=self.device, dtype=torch.float32)
            outs = self._views(static_flat, B)
            s = self._struct(outs, 0, B)
            static_img.copy_(img)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                      # warm-up outside capture (lazy func attributes)
                self.check(self.lib.hm_forward(self.h, _ptr(static_img), B, C.byref(s), self.stream()))
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.check(self.lib.hm_forward(self.h, _ptr(static_img), B, C.byref(s), self.stream()))
            if len(self._graphs) >= self.MAX_GRAPHS:
                torch.cuda.synchronize(self.device)
                self._graphs.pop(next(iter(self._graphs)))      # dicts keep insertion order: first = least recently used
            g = (graph, static_img, static_flat)
        self._graphs[B] = g                                    # (re-)insert at the most-recently-used end
        graph, static_img, static_flat = g
        static_img.copy_(img)
        graph.replay()
        return self._views(static_flat.clone(), B)             # results are the caller's own memory

    def forward(self, img, want_feats=False):
        """img: CUDA float32 [B,3,256,256].  Returns dict of CUDA tensors (hm_outputs fields)."""
        B = img.shape[0]
        if B == 0:                                             # an empty shard (fewer hands than ranks): empty tensors, no launch
            return self._outputs(0, want_feats)
        if (self.use_graphs and not want_feats and not self._profiling and B <= self.max_batch
                and not torch.cuda.is_current_stream_capturing()):
            with torch.cuda.device(self.device):
                return self._forward_graph(img)
        t = self._outputs(B, want_feats)
        with torch.cuda.device(self.device):
            for lo in range(0, B, self.max_batch):          # transparently chunk large batches
                hi = min(B, lo + self.max_batch)
                s = self._struct(t, lo, hi)
                self.check(self.lib.hm_forward(self.h, _ptr(img[lo:hi]), hi - lo, C.byref(s), self.stream()))
        return t

    def backbone(self, img):
        B = img.shape[0]
        feats = torch.empty(B, 192, 1280, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            for lo in range(0, B, self.max_batch):
                hi = min(B, lo + self.max_batch)
                self.check(self.lib.hm_backbone(self.h, _ptr(img[lo:hi]), hi - lo, _ptr(feats[lo:hi]), self.stream()))
        return feats

    def mano_head(self, feats):
        B = feats.shape[0]
        t = self._outputs(B)
        with torch.cuda.device(self.device):
            for lo in range(0, B, self.max_batch):
                hi = min(B, lo + self.max_batch)
                s = self._struct(t, lo, hi)
                self.check(self.lib.hm_mano_head(self.h, _ptr(feats[lo:hi]), hi - lo, C.byref(s), self.stream()))
        return t

    def mano_lbs(self, betas, rotmats, cam_t=None):
        """betas [B,10], rotmats [B,16,3,3] (CUDA f32) -> vertices [B,778,3], joints [B,21,3]
        (+ crop-normalised 2D keypoints [B,21,2] when cam_t [B,3] is given)."""
        B = betas.shape[0]
        verts = torch.empty(B, 778, 3, device=self.device, dtype=torch.float32)
        joints = torch.empty(B, 21, 3, device=self.device, dtype=torch.float32)
        kp2d = torch.empty(B, 21, 2, device=self.device, dtype=torch.float32) if cam_t is not None else None
        with torch.cuda.device(self.device):
            self.check(self.lib.hm_mano_lbs(self.h, _ptr(betas), _ptr(rotmats), _ptr(cam_t), B,
                                            _ptr(verts), _ptr(joints), _ptr(kp2d), self.stream()))
        return (verts, joints, kp2d) if cam_t is not None else (verts, joints)

    def preprocess(self, frame_bgr, boxes_xyxy, is_right, out=None):
        """frame_bgr: HxWx3 uint8 numpy (host) or CUDA uint8 tensor.  Returns
        (img [n,3,256,256] CUDA f32, box_center [n,2], box_size [n], trans [n,2,3]) (numpy f32)."""
        boxes = np.ascontiguousarray(boxes_xyxy, dtype=np.float64).reshape(-1, 4)
        n = boxes.shape[0]
        right = np.ascontiguousarray(is_right, dtype=np.uint8).reshape(-1)
        img = out if out is not None else torch.empty(n, 3, IMG, IMG, device=self.device, dtype=torch.float32)
        center = np.empty((n, 2), np.float32)
        size = np.empty((n,), np.float32)
        trans = np.empty((n, 2, 3), np.float32)
        if isinstance(frame_bgr, _DeviceFrame):
            H, W = frame_bgr.shape[:2]
            fptr, stride, on_dev, keep = frame_bgr.ptr, frame_bgr.row_stride, 1, frame_bgr
        elif isinstance(frame_bgr, torch.Tensor):
            assert frame_bgr.is_cuda and frame_bgr.dtype == torch.uint8 and frame_bgr.stride(-1) == 1
            H, W = frame_bgr.shape[:2]
            fptr, stride, on_dev, keep = frame_bgr.data_ptr(), frame_bgr.stride(0), 1, frame_bgr
        else:
            keep = np.ascontiguousarray(frame_bgr)
            assert keep.dtype == np.uint8 and keep.ndim == 3 and keep.shape[2] == 3
            H, W = keep.shape[:2]
            fptr, stride, on_dev = keep.ctypes.data, keep.strides[0], 0
        with torch.cuda.device(self.device):
            self.check(self.lib.hm_preprocess(
                self.h, C.c_void_p(fptr), H, W, stride, on_dev, C.c_void_p(boxes.ctypes.data),
                C.c_void_p(right.ctypes.data), n, _ptr(img), C.c_void_p(center.ctypes.data),
                C.c_void_p(size.ctypes.data), C.c_void_p(trans.ctypes.data), self.stream()))
        return img, center, size, trans

    def preprocess_centers(self, frame_bgr, center_size, modes, out=None, blur_sigma=None):
        """Crop windows given directly: center_size [n,3] float64 (c_x, c_y, size), modes [n] uint8 (0 plain, 1 mirror the
        patch, 2 warp the mirrored frame - include/hamer_b200.h hm_preprocess_centers).  blur_sigma [n] (0 = none): the
        anti-alias gaussian of ViTDetDataset / prepare_item applied to the frame before the warp (hm_preprocess_blur).
        Returns (img, trans)."""
        ccs = np.ascontiguousarray(center_size, dtype=np.float64).reshape(-1, 3)
        n = ccs.shape[0]
        md = np.ascontiguousarray(modes, dtype=np.uint8).reshape(-1)
        radius = weights = None
        if blur_sigma is not None and any(float(sg) > 0 for sg in blur_sigma):
            radius = np.zeros(n, np.int32)
            weights = np.zeros((n, BLUR_MAX_RADIUS + 1), np.float64)
            for i, sg in enumerate(blur_sigma):
                if float(sg) > 0:
                    radius[i], weights[i] = gaussian_taps(sg)
        img = out if out is not None else torch.empty(n, 3, IMG, IMG, device=self.device, dtype=torch.float32)
        trans = np.empty((n, 2, 3), np.float32)
        if isinstance(frame_bgr, torch.Tensor):
            assert frame_bgr.is_cuda and frame_bgr.dtype == torch.uint8 and frame_bgr.stride(-1) == 1
            H, W = frame_bgr.shape[:2]
            fptr, stride, on_dev, keep = frame_bgr.data_ptr(), frame_bgr.stride(0), 1, frame_bgr
        else:
            keep = np.ascontiguousarray(frame_bgr)
            assert keep.dtype == np.uint8 and keep.ndim == 3 and keep.shape[2] == 3
            H, W = keep.shape[:2]
            fptr, stride, on_dev = keep.ctypes.data, keep.strides[0], 0
        with torch.cuda.device(self.device):
            for lo in range(0, n, self.max_batch):
                hi = min(n, lo + self.max_batch)
                if radius is None:
                    self.check(self.lib.hm_preprocess_centers(
                        self.h, C.c_void_p(fptr), H, W, stride, on_dev, C.c_void_p(ccs[lo:hi].ctypes.data),
                        C.c_void_p(md[lo:hi].ctypes.data), hi - lo, _ptr(img[lo:hi]), C.c_void_p(trans[lo:hi].ctypes.data),
                        self.stream()))
                else:
                    self.check(self.lib.hm_preprocess_blur(
                        self.h, C.c_void_p(fptr), H, W, stride, on_dev, C.c_void_p(ccs[lo:hi].ctypes.data),
                        C.c_void_p(md[lo:hi].ctypes.data), C.c_void_p(radius[lo:hi].ctypes.data),
                        C.c_void_p(weights[lo:hi].ctypes.data), hi - lo, _ptr(img[lo:hi]), C.c_void_p(trans[lo:hi].ctypes.data),
                        self.stream()))
        return img, trans

    def camera_postprocess(self, pred_cam, kp3d, meta, fxfy_differ, add_center):
        """meta: CUDA float32 [n, >=16] (layout: include/hamer_b200.h).  kp3d is flipped in place.
        Returns (cam_t_full [n,3], kp2d_full [n,21,2])."""
        n = pred_cam.shape[0]
        assert meta.is_cuda and meta.dtype == torch.float32 and meta.stride(1) == 1 and kp3d.is_contiguous()
        cam_t = torch.empty(n, 3, device=self.device, dtype=torch.float32)
        kp2d = torch.empty(n, 21, 2, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self.check(self.lib.hm_camera_postprocess(
                self.h, _ptr(pred_cam.contiguous()), _ptr(kp3d), _ptr(meta), int(meta.stride(0)), int(bool(fxfy_differ)),
                int(bool(add_center)), n, _ptr(cam_t), _ptr(kp2d), self.stream()))
        return cam_t, kp2d

    def rotmat_to_axis_angle(self, rotmats):
        """[..., 3, 3] CUDA float32 rotation matrices -> [..., 3] axis-angle vectors (cv2.Rodrigues' log map on the device)."""
        r = rotmats.to(self.device, torch.float32).contiguous()
        n = r.numel() // 9
        out = torch.empty(*r.shape[:-2], 3, device=self.device, dtype=torch.float32)
        if n:
            with torch.cuda.device(self.device):
                self.check(self.lib.hm_rotmat_to_axis_angle(self.h, _ptr(r), n, _ptr(out), self.stream()))
        return out

    def letterbox(self, frame_bgr, new_size=640, stride=32, half=False):
        """LoadImage.process_img + the / 255 of Detector.detect on the device (hm_letterbox): HxWx3 uint8 BGR frame (numpy on the
        host or a CUDA uint8 tensor) -> ([3, H', W'] CUDA float32 / float16 RGB in 0..1, the frame on the device as a CUDA
        uint8 view).  A host frame is uploaded once; pass the returned device frame to ``preprocess`` for the crops."""
        oh, ow = C.c_int(), C.c_int()
        if isinstance(frame_bgr, torch.Tensor):
            assert frame_bgr.is_cuda and frame_bgr.dtype == torch.uint8 and frame_bgr.stride(-1) == 1
            H, W = frame_bgr.shape[:2]
            fptr, rstride, on_dev, keep = frame_bgr.data_ptr(), frame_bgr.stride(0), 1, frame_bgr
        else:
            keep = np.ascontiguousarray(frame_bgr)
            assert keep.dtype == np.uint8 and keep.ndim == 3 and keep.shape[2] == 3
            H, W = keep.shape[:2]
            fptr, rstride, on_dev = keep.ctypes.data, keep.strides[0], 0
        self.check(self.lib.hm_letterbox_shape(H, W, int(new_size), int(stride), C.byref(oh), C.byref(ow)))
        img = torch.empty(3, oh.value, ow.value, device=self.device, dtype=torch.float16 if half else torch.float32)
        dev = C.c_void_p()
        with torch.cuda.device(self.device):
            self.check(self.lib.hm_letterbox(self.h, C.c_void_p(fptr), H, W, rstride, on_dev, int(new_size), int(stride),
                                             int(bool(half)), _ptr(img), C.byref(dev), self.stream()))
        if on_dev:
            return img, frame_bgr
        return img, _DeviceFrame(dev.value, H, W, rstride, keep)

    def scale_boxes(self, det, img1_shape, img0_shape):
        """scale_coords(...).round() + labels (hm_scale_boxes): det CUDA [n, >=6] float32 / float16 (x1 y1 x2 y2 conf cls) ->
        the reference's detection list [[label, [x1, y1, x2, y2]], ...] (one small device -> host copy)."""
        n = det.shape[0]
        if n == 0:
            return []
        assert det.is_cuda and det.dtype in (torch.float32, torch.float16) and det.stride(1) == 1 and det.shape[1] >= 6
        boxes = torch.empty(n, 4, device=self.device, dtype=torch.float64)
        right = torch.empty(n, device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            self.check(self.lib.hm_scale_boxes(self.h, _ptr(det), int(det.dtype == torch.float16), n, int(det.stride(0)),
                                               int(img1_shape[0]), int(img1_shape[1]), int(img0_shape[0]), int(img0_shape[1]),
                                               _ptr(boxes), _ptr(right), self.stream()))
        b, r = boxes.cpu().tolist(), right.cpu().tolist()
        return [['right' if r[i] else 'left', b[i]] for i in range(n)]

    def launch_count(self):
        return int(self.lib.hm_forward_launch_count(self.h))

    def profile_enable(self, on=True):
        self._profiling = bool(on)                             # per-launch events cannot be captured in a graph
        self.check(self.lib.hm_profile_enable(self.h, 1 if on else 0))

    def profile_collect(self):
        """-> {category: (total_ms, spans)} since the last collect (synchronises the device)."""
        n = len(_lib.PROFILE_CATEGORIES)
        ms = (C.c_double * n)()
        cnt = (C.c_int32 * n)()
        self.check(self.lib.hm_profile_collect(self.h, ms, cnt, n))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_lib.PROFILE_CATEGORIES)}


class HAMER:
    """The reference's ``HAMER`` (inference surface).  ``model(batch)`` returns the same two
    dicts as hamer.py:124-156; tensors live on the model's CUDA device, float32."""

    def __init__(self, cfg=None, init_renderer=False, mano=None, device='cuda:0', max_batch=64,
                 max_frame_bytes=1920 * 1080 * 3, gemm_cta_group=0, use_graphs=True, deterministic=False,
                 split_weights=False):
        self.cfg = cfg if cfg is not None else default_config()
        if self.cfg.MODEL.IMAGE_SIZE != 256:
            raise AssertionError(f'MODEL.IMAGE_SIZE ({self.cfg.MODEL.IMAGE_SIZE}) should be 256 for ViT backbone')
        self.engine = Engine(device, max_batch, max_frame_bytes, gemm_cta_group, use_graphs=use_graphs,
                             deterministic=int(deterministic), split_weights=int(split_weights))
        self.device = self.engine.device
        self.training = False
        if mano is None:
            mano = self._load_mano_from_cfg()
        self.mano = _ManoHandle(mano)
        self.engine.set_mano(mano)
        self._have_weights = False

    def _load_mano_from_cfg(self):
        path = os.path.join(str(self.cfg.MANO.MODEL_PATH), 'MANO_RIGHT.pkl')
        if not os.path.exists(path):
            raise FileNotFoundError(f'MANO model not found at {path} (cfg.MANO.MODEL_PATH); pass mano=... explicitly')
        return load_mano_pkl(path)

    # -- nn.Module-like surface the reference's callers use --------------------------------------
    def to(self, device):
        d = torch.device(device)
        if d.type != 'cuda':
            raise RuntimeError('hamer_b200.HAMER has no CPU path')
        idx = d.index if d.index is not None else torch.cuda.current_device()
        if idx != self.engine.index:
            raise RuntimeError(f'model is bound to cuda:{self.engine.index}; build it with device=cuda:{idx}')
        return self

    def eval(self):
        return self

    def load_state_dict(self, state_dict, strict=False):
        """Accepts the reference checkpoint's key layout; other keys (e.g. the
        discriminator, 'mano.*' buffers) are ignored like load_from_checkpoint(strict=False)."""
        used = 0
        for k, v in state_dict.items():
            if k.startswith('backbone.') or k.startswith('mano_head.'):
                self.engine.set_weight(k, v)
                used += 1
        self.engine.finalize_weights()
        self._have_weights = True
        return used

    def forward_step(self, batch, train=False):
        if train:
            raise NotImplementedError('inference only')
        x = batch['img']
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, IMG, IMG):
            raise ValueError(f"batch['img'] must be [B,3,256,256], got {tuple(x.shape)}")
        x = x.to(self.device, torch.float32).contiguous()
        B = x.shape[0]
        t = self.engine.forward(x, want_feats=False)
        pred_mano_params = {'global_orient': t['global_orient'], 'hand_pose': t['hand_pose'], 'betas': t['betas']}
        output = {
            'pred_cam': t['pred_cam'],
            'pred_mano_params': {k: v.clone() for k, v in pred_mano_params.items()},   # hamer.py:126
            'pred_cam_t': t['pred_cam_t'],
            'focal_length': t['focal_length'],
            'pred_keypoints_3d': t['pred_keypoints_3d'].reshape(B, 21, 3),
            'pred_vertices': t['pred_vertices'].reshape(B, 778, 3),
            'pred_keypoints_2d': t['pred_keypoints_2d'].reshape(B, 21, 2),
        }
        pred_mano_params['trans'] = t['pred_cam_t']                                     # hamer.py:155
        return output, pred_mano_params

    def forward(self, batch):
        return self.forward_step(batch, train=False)

    __call__ = forward


class HAMER_INFER(HAMER):
    """The reference's inference-only module (hamer/hamer/models/hamer.py:468-563): ``forward(batch)`` returns ONE dict -
    the ``output`` of HAMER.forward_step plus 'pose' (B,16,3,3) = cat(global_orient, hand_pose) and 'beta' (B,10)
    (hamer.py:536-541) - and no second pred_mano_params dict."""

    def forward_step(self, batch, train=False):
        output, params = super().forward_step(batch, train=train)
        output['pose'] = torch.cat([params['global_orient'], params['hand_pose']], dim=1)      # hamer.py:538-540
        output['beta'] = params['betas']                                                      # hamer.py:541
        return output

    def forward(self, batch):
        return self.forward_step(batch, train=False)

    __call__ = forward


def load_hamer(checkpoint_path=None, mano=None, device='cuda:0', max_batch=64, cfg=None, state_dict=None,
               deterministic=False, split_weights=False):
    """load_hamer (models/__init__.py:32-47): returns (model, model_cfg).  ``checkpoint_path`` is a
    Lightning checkpoint ({'state_dict': ...}) or a plain state dict saved with torch.save;
    ``state_dict`` can be passed directly instead.  ``deterministic=True`` fixes the fp32 summation order of the
    residual GEMMs (bit-identical results across runs and batch sizes, a few percent slower).
    ``split_weights=True`` keeps the backbone weights as fp16 hi + lo pairs (fp32-faithful weights for checkpoints whose
    values are not fp16-representable - a trained HaMeR checkpoint - at twice the tensor-core work)."""
    cfg = cfg if cfg is not None else default_config()
    if cfg.MODEL.BACKBONE.TYPE == 'vit' and 'BBOX_SHAPE' not in cfg.MODEL:
        assert cfg.MODEL.IMAGE_SIZE == 256
        cfg.MODEL.BBOX_SHAPE = [192, 256]
    model = HAMER(cfg, mano=mano, device=device, max_batch=max_batch, deterministic=deterministic,
                  split_weights=split_weights)
    if state_dict is None:
        if checkpoint_path is None:
            raise ValueError('load_hamer needs checkpoint_path or state_dict')
        ckpt = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
        state_dict = ckpt.get('state_dict', ckpt)
    model.load_state_dict(state_dict, strict=False)
    return model, cfg
