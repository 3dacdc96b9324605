"""Geodesic / quaternion-L1 rotation loss + L1 translation loss on B200 -- drop-in for
the reference's ``models/pose_loss.py`` (SFR-Vision/6d-pose-estimation).

``PoseLoss.forward`` returns the same 0-d tensor as the reference and supports
``.backward()``; the forward value and the gradients w.r.t. the predictions come out of
ONE kernel launch (``p6d_pose_loss_fwd_bwd``) instead of ~40 eager launches
(reference ``pose_loss.py:19-61`` + autograd).  CUDA only, no CPU path.
"""
import os

import torch
import torch.nn as nn


try:                                    # imported as part of the package ...
    from .._p6d_bootstrap import core as _core
except ImportError:                     # ... or as top-level `models` / `utils` (drop-in layout: this
    from _p6d_bootstrap import core as _core   # directory is at the front of sys.path, see dropin.py)


_workspaces = {}


def _workspace(dev):
    ws = _workspaces.get(dev)
    if ws is None:
        ws = torch.zeros(int(_core().lib().p6d_pose_loss_workspace_bytes()), dtype=torch.uint8, device=dev)
        _workspaces[dev] = ws
    return ws


def _ready(t, dev, tail):
    """`t` as a contiguous float32 tensor on `dev` -- without touching it when it already is one
    (the training step: every input comes straight out of the network; each avoided torch op is
    3-5 us of the ~100 us the Python side of a B = 32 step costs)."""
    if (isinstance(t, torch.Tensor) and t.dtype is torch.float32 and t.device == dev and t.is_contiguous()
            and t.data_ptr() % 16 == 0):
        return t          # only its pointer and size are used below: no detach needed
    return _core().as_cuda_f32(t, dev, tail)


def _row_major(shape):
    """Contiguous strides of `shape`."""
    strides, step = [], 1
    for n in reversed(shape):
        strides.append(step)
        step *= max(int(n), 1)
    return tuple(reversed(strides))


def _empty_batch_nan(*tensors):
    """The reference on an empty batch: mean over zero rows -> NaN, attached to the inputs' graph."""
    total = None
    for t in tensors:
        s = t.sum() * float("nan")
        total = s if total is None else total + s
    return total


class _FusedPoseLoss(torch.autograd.Function):
    """out = [loss, rot_term, trans_term]; grads for upstream 1 are produced by the
    forward launch and scaled by grad_output in backward."""

    @staticmethod
    def forward(ctx, pred_rot, pred_trans, gt_rot, gt_trans, rot_weight, trans_weight, mode, pick):
        core = _core()
        dev = core.require_cuda(pred_rot.device)
        pq = _ready(pred_rot, dev, (4,))
        pt = _ready(pred_trans, dev, (3,))
        gq = _ready(gt_rot, dev, (4,))
        gt = _ready(gt_trans, dev, (3,))
        B = pq.numel() // 4
        if not (pt.numel() == 3 * B and gq.numel() == 4 * B and gt.numel() == 3 * B) or B == 0:
            raise ValueError("PoseLoss needs inputs with a common batch dimension")
        need_q = ctx.needs_input_grad[0]
        need_t = ctx.needs_input_grad[1]
        # one allocation: [grad_q 4B | grad_t 3B | pad to 4 | loss, rot term, trans term]; both gradients
        # in one buffer so that backward scales them with one launch
        gb = (7 * B + 3) // 4 * 4
        buf = torch.empty(gb + 4, dtype=torch.float32, device=dev)
        base = buf.data_ptr()
        core.check(core.lib().p6d_pose_loss_fwd_bwd(
            pq.data_ptr(), pt.data_ptr(), gq.data_ptr(), gt.data_ptr(), B, float(rot_weight), float(trans_weight),
            int(mode), base + 4 * gb, base if need_q else None, base + 16 * B if need_t else None,
            _workspace(dev).data_ptr(), dev.index, core.stream_ptr(dev)))
        ctx.buf = buf
        ctx.meta = (pred_rot.shape, pred_trans.shape, pred_rot.dtype, pred_trans.dtype, need_q, need_t, B)
        return buf[gb + pick]

    @staticmethod
    def backward(ctx, grad_out):
        rs, ts, rd, td, need_q, need_t, B = ctx.meta
        # pick 0: d loss; pick 1: d rot_term (the kernel ran with rot_weight 1, trans_weight 0)
        # three torch ops instead of six: the whole buffer scaled at once (its tail -- padding and the loss
        # terms -- is never looked at), each gradient one as_strided view of the product
        scaled = ctx.buf * grad_out
        dq = dt = None
        if need_q:
            dq = scaled.as_strided(rs, _row_major(rs), 0)
            dq = dq if rd is torch.float32 else dq.to(rd)
        if need_t:
            dt = scaled.as_strided(ts, _row_major(ts), 4 * B)
            dt = dt if td is torch.float32 else dt.to(td)
        return dq, dt, None, None, None, None, None, None


class _FusedGeometricPoseLoss(torch.autograd.Function):
    """PoseLoss on a pinhole translation computed in the same launch (kernels d1 + c):
    returns (loss, translation); gradients flow to pred_rot and z_pred."""

    @staticmethod
    def forward(ctx, pred_rot, z_pred, bbox_center, camera_matrix, gt_rot, gt_trans, rot_weight, trans_weight, mode):
        core = _core()
        dev = core.require_cuda(pred_rot.device)
        pq = _ready(pred_rot, dev, (4,))
        z = _ready(z_pred, dev, ())
        uv = _ready(bbox_center, dev, (2,))
        gq = _ready(gt_rot, dev, (4,))
        gt = _ready(gt_trans, dev, (3,))
        B = pq.numel() // 4
        if B == 0 or not (z.numel() == B and uv.numel() == 2 * B and gq.numel() == 4 * B and gt.numel() == 3 * B):
            raise ValueError("forward_geometric needs non-empty inputs with a common batch dimension")
        K = _ready(camera_matrix, dev, ())
        if camera_matrix.dim() == 2 and K.numel() == 9:
            kb = 0
        elif camera_matrix.dim() == 3 and K.numel() == 9 * B:
            kb = 1
        else:
            raise ValueError("camera_matrix must be [3,3] or [B,3,3]")
        need_q, need_z = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        # one allocation: [grad_q 4B | grad_z B | pad to 4 | translation 3B | pad to 4 | loss, rot, trans terms]
        gb = (5 * B + 3) // 4 * 4
        tb = (3 * B + 3) // 4 * 4
        buf = torch.empty(gb + tb + 4, dtype=torch.float32, device=dev)
        base = buf.data_ptr()
        core.check(core.lib().p6d_pose_loss_pinhole_fwd_bwd(
            pq.data_ptr(), z.data_ptr(), uv.data_ptr(), K.data_ptr(), kb, gq.data_ptr(), gt.data_ptr(), B,
            float(rot_weight), float(trans_weight), int(mode), base + 4 * (gb + tb), base if need_q else None,
            base + 16 * B if need_z else None, base + 4 * gb, _workspace(dev).data_ptr(), dev.index,
            core.stream_ptr(dev)))
        trans = buf[gb:gb + 3 * B].view(B, 3)
        ctx.buf = buf
        ctx.meta = (pred_rot.shape, z_pred.shape, pred_rot.dtype, z_pred.dtype, need_q, need_z, B)
        ctx.mark_non_differentiable(trans)
        return buf[gb + tb], trans

    @staticmethod
    def backward(ctx, grad_loss, _grad_trans):
        rs, zs, rd, zd, need_q, need_z, B = ctx.meta
        scaled = ctx.buf * grad_loss
        dq = dz = None
        if need_q:
            dq = scaled.as_strided(rs, _row_major(rs), 0)
            dq = dq if rd is torch.float32 else dq.to(rd)
        if need_z:
            dz = scaled.as_strided(zs, _row_major(zs), 4 * B)
            dz = dz if zd is torch.float32 else dz.to(zd)
        return dq, dz, None, None, None, None, None, None, None


class CapturedPoseLossStep:
    """The loss step of a training iteration -- ``criterion(...)`` + the gradients w.r.t. the
    predictions -- captured ONCE in a CUDA graph and replayed with a single launch (addition to the
    reference surface; see ``PoseLoss.capture``).

    At the reference's batch size (32) the kernel needs 7-9 us while the eager call costs ~100 us of
    Python / autograd bookkeeping per step; a replay costs one ``cudaGraphLaunch``.  The tensors are
    STATIC: write the network outputs into ``pred_rot`` / ``pred_trans`` (or ``z_pred`` /
    ``bbox_center`` / ``camera_matrix`` for the geometric form) and the targets into ``gt_rot`` /
    ``gt_trans`` (``copy_``, or let a captured network produce them in place), call ``replay()``, read
    ``loss``, ``grad_rot`` and ``grad_trans`` (``grad_z`` for the geometric form).  ``__call__`` takes
    the caller's own tensors: ready ones (float32, contiguous, on the device) are read in place by one
    direct launch that writes the same static outputs; anything else is copied into the static
    inputs (one fused ``_foreach_copy_``) and the graph replayed."""

    def __init__(self, criterion, pred_rot, pred_trans, gt_rot, gt_trans, geometric=None, warmup=3):
        dev = _core().require_cuda(pred_rot.device)
        static = lambda t, grad=False: t.detach().to(dev, torch.float32).clone().contiguous().requires_grad_(grad)
        self.pred_rot = static(pred_rot, True)
        self.gt_rot, self.gt_trans = static(gt_rot), static(gt_trans)
        self.geometric = geometric is not None
        self._weights = (float(criterion.rot_weight), float(criterion.trans_weight), int(criterion._mode()))
        if self.geometric:            # pred_trans is z_pred; geometric = (bbox_center, camera_matrix)
            self.z_pred = static(pred_trans, True)
            self.bbox_center, self.camera_matrix = static(geometric[0]), static(geometric[1])
            leaves = (self.pred_rot, self.z_pred)
            run = lambda: criterion.forward_geometric(self.pred_rot, self.z_pred, self.bbox_center, self.camera_matrix,
                                                      self.gt_rot, self.gt_trans)
            self._ins = [self.pred_rot, self.z_pred, self.bbox_center, self.camera_matrix, self.gt_rot, self.gt_trans]
        else:
            self.pred_trans = static(pred_trans, True)
            leaves = (self.pred_rot, self.pred_trans)
            run = lambda: (criterion(self.pred_rot, self.pred_trans, self.gt_rot, self.gt_trans), None)
            self._ins = [self.pred_rot, self.pred_trans, self.gt_rot, self.gt_trans]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):               # warm-up outside the capture (allocator, lazy loading)
            for _ in range(max(1, warmup)):
                torch.autograd.grad(run()[0], leaves)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.translation = run()
            g = torch.autograd.grad(self.loss, leaves)
        self.grad_rot = g[0]
        if self.geometric:
            self.grad_z = g[1]
        else:
            self.grad_trans = g[1]

    def replay(self):
        self.graph.replay()
        return self.loss

    def _launch_on(self, tensors):
        """The captured step is ONE kernel (the backward's scale by 1 changes no bit), so for inputs that
        are already what the kernel reads -- float32, contiguous, 16-byte aligned, on the device, of the
        captured sizes -- it is launched straight on the caller's tensors, with the static ``loss`` /
        ``grad_*`` / ``translation`` tensors as its outputs: no copies, no graph, ~one ctypes call.
        Returns False when an input needs converting (the caller then copies and replays)."""
        dev = self.pred_rot.device
        for t, s in zip(tensors, self._ins):
            if not (isinstance(t, torch.Tensor) and t.dtype is torch.float32 and t.device == dev and t.is_contiguous()
                    and t.data_ptr() % 16 == 0 and t.numel() == s.numel()):
                return False
        core = _core()
        B = self.pred_rot.numel() // 4
        ws = _workspace(dev).data_ptr()
        if self.geometric:
            pq, z, uv, K, gq, gt = tensors
            core.check(core.lib().p6d_pose_loss_pinhole_fwd_bwd(
                pq.data_ptr(), z.data_ptr(), uv.data_ptr(), K.data_ptr(), 1 if self.camera_matrix.dim() == 3 else 0,
                gq.data_ptr(), gt.data_ptr(), B, *self._weights, self.loss.data_ptr(), self.grad_rot.data_ptr(),
                self.grad_z.data_ptr(), self.translation.data_ptr(), ws, dev.index, core.stream_ptr(dev)))
        else:
            pq, pt, gq, gt = tensors
            core.check(core.lib().p6d_pose_loss_fwd_bwd(
                pq.data_ptr(), pt.data_ptr(), gq.data_ptr(), gt.data_ptr(), B, *self._weights, self.loss.data_ptr(),
                self.grad_rot.data_ptr(), self.grad_trans.data_ptr(), ws, dev.index, core.stream_ptr(dev)))
        return True

    def __call__(self, *tensors):
        if len(tensors) != len(self._ins):
            raise ValueError(f"expected {len(self._ins)} tensors")
        return self.loss if self._launch_on(tensors) else self.copy_and_replay(*tensors)

    def copy_and_replay(self, *tensors):
        with torch.no_grad():
            torch._foreach_copy_(self._ins, [t.detach() for t in tensors])
        return self.replay()


class PoseLoss(nn.Module):
    """rot_weight * rotation_loss + trans_weight * L1(translation)  (reference :8-65)."""

    def __init__(self, rot_weight=1.0, trans_weight=1.0, rotation_loss='geodesic'):
        super().__init__()
        self.rot_weight = rot_weight
        self.trans_weight = trans_weight
        self.rotation_loss_type = rotation_loss

    def _mode(self):
        # anything other than 'geodesic' selects the quaternion-L1 distance (reference :21-24)
        return 0 if self.rotation_loss_type == 'geodesic' else 1

    def forward(self, pred_rot, pred_trans, gt_rot, gt_trans, obj_ids=None):
        if pred_rot.shape[0] == 0:
            return _empty_batch_nan(pred_rot, pred_trans)
        return _FusedPoseLoss.apply(pred_rot, pred_trans, gt_rot, gt_trans, self.rot_weight,
                                    self.trans_weight, self._mode(), 0)

    def forward_geometric(self, pred_rot, z_pred, bbox_center, camera_matrix, gt_rot, gt_trans):
        """The RGB-Geometric training step in one launch (addition to the reference surface):
        equals ``self(pred_rot, pinhole_translation(z_pred, bbox_center, camera_matrix), gt_rot,
        gt_trans)`` bit for bit -- the model's pinhole translation
        (reference models/pose_net_rgb_geometric.py:93-109) is computed inside the loss kernel
        and the gradient comes back w.r.t. ``z_pred``.  Returns (loss, translation [B,3])."""
        return _FusedGeometricPoseLoss.apply(pred_rot, z_pred, bbox_center, camera_matrix, gt_rot, gt_trans,
                                             self.rot_weight, self.trans_weight, self._mode())

    def capture(self, pred_rot, pred_trans, gt_rot, gt_trans, geometric=None):
        """CUDA-graph form of the training step for fixed shapes: see ``CapturedPoseLossStep``.
        ``geometric=(bbox_center, camera_matrix)`` captures ``forward_geometric`` with
        ``pred_trans`` taken as ``z_pred``."""
        return CapturedPoseLossStep(self, pred_rot, pred_trans, gt_rot, gt_trans, geometric)

    def _rotation_only(self, q1, q2, mode):
        zeros = torch.zeros(q1.shape[0], 3, dtype=torch.float32, device=q1.device)
        return _FusedPoseLoss.apply(q1, zeros, q2, zeros, 1.0, 0.0, mode, 1)

    def _geodesic_distance(self, q1, q2):
        """mean_b 2*atan2(|q1-q2'|, |q1+q2'|) on normalised quaternions (reference :30-50)."""
        return self._rotation_only(q1, q2, 0)

    def _quaternion_l1(self, q1, q2):
        """mean_b min(sum|q1-q2|, sum|q1+q2|) on normalised quaternions (reference :52-61)."""
        return self._rotation_only(q1, q2, 1)

    def train_loss(self, pred_rot, pred_trans, gt_rot, gt_trans, obj_ids=None):
        return self.forward(pred_rot, pred_trans, gt_rot, gt_trans, obj_ids)
