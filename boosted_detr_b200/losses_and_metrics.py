"""Hungarian set-matching loss — same classes / call conventions as the reference's
ModelComponents/losses_and_metrics.py, backed by the sm_100a kernels in csrc/matcher.cu.

Tensors are CUDA torch tensors used as plain device buffers (fp32; num_objects int32).  No CPU
fallback: every call goes through libbdetr.so.
"""
from __future__ import annotations

import torch

from . import _lib
from .device import empty, f32, i32, ptr, stream_ptr, zeros

# default weights (reference losses_and_metrics.py:8-11)
DEFAULT_CATEGORY_WEIGHT = 1000.0
DEFAULT_BOX_WEIGHT = 1.0
DEFAULT_ATTRIBUTE_WEIGHT = 100.0
DEFAULT_EXIST_WEIGHT = 100.0


# The reference's loss functions (:44-72) are elementwise over broadcast shapes and are only ever evaluated through
# CostArray.call(y_true, y_pred, func), which builds [B,T,1,K] x [B,1,Q,K] operands.  Here they are callable markers:
# CostArray selects the term by identity, and a direct call evaluates the same pairwise [B,T,Q] cost on the device --
# with either the reference's broadcast operands ([B,T,1,K], [B,1,Q,K]) or the plain ([B,T,K], [B,Q,K]) pair.
def _pairwise_call(func, y_true, y_pred):
    yt, yp = f32(y_true), f32(y_pred)
    if yt.dim() == 4 and yt.shape[2] == 1:
        yt = yt[:, :, 0, :].contiguous()
    if yp.dim() == 4 and yp.shape[1] == 1:
        yp = yp[:, 0, :, :].contiguous()
    if yt.dim() != 3 or yp.dim() != 3:
        raise ValueError("expected [B,T,K] / [B,Q,K] (or the reference's [B,T,1,K] / [B,1,Q,K]) operands")
    return CostArray().call(yt, yp, func)


def CategoryLoss(y_true, y_pred):  # reference :44-49
    return _pairwise_call(CategoryLoss, y_true, y_pred)


def AttributeLoss(y_true, y_pred):  # reference :51-57
    return _pairwise_call(AttributeLoss, y_true, y_pred)


def BoxLoss(y_true, y_pred):  # reference :68-72
    return _pairwise_call(BoxLoss, y_true, y_pred)


def raise_for_status(status: torch.Tensor) -> None:
    """Mirrors scipy's ValueErrors (the reference dies on them inside tf.numpy_function)."""
    st = status.cpu()
    if (st == _lib.BDETR_E_INVALID_COST).any():
        raise ValueError("matrix contains invalid numeric entries")
    if (st == _lib.BDETR_E_INFEASIBLE).any():
        raise ValueError("cost matrix is infeasible")


class PreparedTargets:
    """Target side of the cost matrix, digested once (bdetr_cost_targets_prepare) and reusable against any number of
    prediction sets of the same batch -- the boosted model matches ONE target batch at every block."""

    def __init__(self, y_true):
        category, attribute, bbox = (f32(t) for t in y_true[:3])
        self.B, self.T, self.C = category.shape
        self.A = attribute.shape[2]
        self.true = (category, attribute, bbox)
        nbytes = _lib.load().bdetr_cost_targets_bytes(self.B, self.T, self.C, self.A)
        self.buffer = torch.empty(nbytes, dtype=torch.uint8, device=category.device)
        _lib.call("bdetr_cost_targets_prepare", self.B, self.T, self.C, self.A, ptr(category), ptr(attribute), ptr(bbox),
                  ptr(self.buffer), stream_ptr())


def pairwise_cost(y_true, y_pred, w_cat, w_box, w_attr, prepared=None):
    """Weighted [B,T,Q] matching cost (reference MatchingLoss.call :119-130)."""
    cat_preds, attr_preds, box_preds = (f32(t) for t in y_pred)
    if prepared is None:
        prepared = PreparedTargets(y_true)
    B, T, C, A = prepared.B, prepared.T, prepared.C, prepared.A
    Q = cat_preds.shape[1]
    cost = empty(B, T, Q)
    _lib.call("bdetr_cost_matrix_prepared", B, T, Q, C, A, ptr(prepared.buffer),
              ptr(cat_preds), ptr(attr_preds), ptr(box_preds), float(w_cat), float(w_box), float(w_attr),
              ptr(cost), stream_ptr())
    return cost


class CostArray:
    """Pairwise f(x, y) values, targets on rows and predictions on columns (reference :215-225)."""

    def __init__(self, **kwargs):
        self.name = "CostArray"

    def call(self, y_true, y_pred, func):
        y_true, y_pred = f32(y_true), f32(y_pred)
        B, T, Q = y_true.shape[0], y_true.shape[1], y_pred.shape[1]
        one = lambda n, k: zeros(B, n, k)
        if func is CategoryLoss:
            tr = [y_true, one(T, 2), one(T, 4)]
            pr = [y_pred, one(Q, 2) + 0.5, one(Q, 4)]
            return pairwise_cost(tr, pr, 1.0, 0.0, 0.0)
        if func is AttributeLoss:
            tr = [one(T, 2), y_true, one(T, 4)]
            pr = [one(Q, 2) + 0.5, y_pred, one(Q, 4)]
            return pairwise_cost(tr, pr, 0.0, 0.0, 1.0)
        if func is BoxLoss:
            tr = [one(T, 2), one(T, 2), y_true]
            pr = [one(Q, 2) + 0.5, one(Q, 2) + 0.5, y_pred]
            return pairwise_cost(tr, pr, 0.0, 1.0, 0.0)
        raise ValueError("func must be CategoryLoss, AttributeLoss or BoxLoss")

    __call__ = call


class MatchingAssignment:
    """Bipartite assignment; bit-identical to scipy.optimize.linear_sum_assignment (reference :228-251)."""

    def __init__(self, name="MatchingAssignment", **kwargs):
        self.name = name
        self.last_status = None

    def assign(self, cost_array, num_objects, want_mask=True, want_assigned=True):
        cost_array = f32(cost_array)
        B, T, Q = cost_array.shape
        num_objects = i32(num_objects.reshape(-1))
        col4row = empty(B, T, dtype=torch.int32)
        row4col = empty(B, Q, dtype=torch.int32)
        status = empty(B, dtype=torch.int32)
        mask = empty(B, T, Q) if want_mask else None
        assigned = empty(B, Q) if want_assigned else None
        _lib.call("bdetr_lsap_assign", B, T, Q, ptr(cost_array), ptr(num_objects), ptr(col4row), ptr(row4col),
                  ptr(mask), ptr(assigned), ptr(status), stream_ptr())
        self.last_status = status
        return col4row, row4col, mask, assigned, status

    def call(self, cost_array, num_objects):
        _, _, mask, _, status = self.assign(cost_array, num_objects, want_assigned=False)
        raise_for_status(status)
        return mask

    __call__ = call


class MatchingMask:
    """(mask [B,T,Q], assigned_predictions [B,Q,1]) (reference :195-212)."""

    def __init__(self, name="MatchingMask", **kwargs):
        self.name = name
        self.MatchingAssignment = MatchingAssignment()

    def call(self, inputs):
        matching_costs, num_objects = inputs
        _, _, mask, assigned, status = self.MatchingAssignment.assign(matching_costs, num_objects)
        raise_for_status(status)
        return mask, assigned.unsqueeze(-1)

    __call__ = call


class MatchingLoss:
    """reference :75-161.  call([y_true, y_pred]) -> ([total, cat, attr, box, exist] each [B], [iou [1,Q]])."""

    def __init__(self, name="MatchingLoss", category_weight=None, box_weight=None, attribute_weight=None,
                 exist_weight=None, **kwargs):
        self.name = name
        self.category_weight = DEFAULT_CATEGORY_WEIGHT if category_weight is None else float(category_weight)
        self.box_weight = DEFAULT_BOX_WEIGHT if box_weight is None else float(box_weight)
        self.attribute_weight = DEFAULT_ATTRIBUTE_WEIGHT if attribute_weight is None else float(attribute_weight)
        self.exist_weight = DEFAULT_EXIST_WEIGHT if exist_weight is None else float(exist_weight)
        self.MatchingMask = MatchingMask()
        self.CostArray = CostArray()
        self.MatchingMetric = MatchingMetric()
        self.check_status = True      # poll the device status flag after each call (set False inside graphs)

    def forward(self, y_true, y_pred, prepared=None):
        """Device-only forward; returns a context dict for `backward` (no host sync).  `prepared` = PreparedTargets of
        y_true when the caller matches the same targets several times."""
        category, attribute, bbox = (f32(t) for t in y_true[:3])
        num_objects = i32(y_true[3].reshape(-1))
        cat_preds, attr_preds, box_preds = (f32(t) for t in y_pred)
        B, T, C = category.shape
        Q, A = cat_preds.shape[1], attribute.shape[2]
        w = (self.category_weight, self.box_weight, self.attribute_weight, self.exist_weight)
        cost = pairwise_cost([category, attribute, bbox], [cat_preds, attr_preds, box_preds], w[0], w[1], w[2], prepared)
        col4row, row4col, _, _, status = self.MatchingMask.MatchingAssignment.assign(
            cost, num_objects, want_mask=False, want_assigned=False)
        losses = empty(5, B)
        iou = empty(Q)
        st = stream_ptr()
        _lib.call("bdetr_matched_loss_fwd", B, T, Q, C, A, ptr(category), ptr(attribute), ptr(bbox), ptr(num_objects),
                  ptr(cat_preds), ptr(attr_preds), ptr(box_preds), ptr(col4row), ptr(row4col),
                  w[0], w[1], w[2], w[3], ptr(losses), ptr(iou), st)
        return {"dims": (B, T, Q, C, A), "true": (category, attribute, bbox, num_objects),
                "pred": (cat_preds, attr_preds, box_preds), "col4row": col4row, "row4col": row4col,
                "status": status, "losses": losses, "iou": iou, "cost": cost, "weights": w}

    def backward(self, ctx, d_cat, d_attr, d_box, gscale=1.0):
        """Accumulates d(gscale * sum_b total_b)/d(y_pred) into d_cat/d_attr/d_box."""
        B, T, Q, C, A = ctx["dims"]
        category, attribute, bbox, num_objects = ctx["true"]
        cat_preds, attr_preds, box_preds = ctx["pred"]
        w = ctx["weights"]
        _lib.call("bdetr_matched_loss_bwd", B, T, Q, C, A, ptr(category), ptr(attribute), ptr(bbox), ptr(num_objects),
                  ptr(cat_preds), ptr(attr_preds), ptr(box_preds), ptr(ctx["col4row"]), ptr(ctx["row4col"]),
                  w[0], w[1], w[2], w[3], float(gscale), ptr(d_cat), ptr(d_attr), ptr(d_box), stream_ptr())

    def call(self, inputs):
        y_true, y_pred = inputs
        ctx = self.forward(y_true, y_pred)
        if self.check_status:
            raise_for_status(ctx["status"])
        self.last_ctx = ctx
        L = ctx["losses"]
        return [L[0], L[1], L[2], L[3], L[4]], [ctx["iou"].unsqueeze(0)]

    __call__ = call


class MatchingMetric:
    """Masked pairwise IoU [B,T,Q] (reference :164-192)."""

    def __init__(self, name="MatchingMetric", **kwargs):
        self.name = name

    def call(self, inputs, assignment_mask=None):
        """inputs = [bbox_true [B,T,4], bbox_pred [B,Q,4]] (COCO x,y,w,h) -> mask * pairwise IoU [B,T,Q] (reference :187-188).
        MatchingLoss.call produces the reduced `IOU` metric itself (fused into the matched-loss kernel)."""
        box_true, box_pred = (f32(t) for t in inputs)
        B, T, Q = box_true.shape[0], box_true.shape[1], box_pred.shape[1]
        out = empty(B, T, Q)
        m = None if assignment_mask is None else f32(assignment_mask)
        _lib.call("bdetr_pairwise_iou", B, T, Q, ptr(box_true), ptr(box_pred), ptr(m), ptr(out), stream_ptr())
        return out

    __call__ = call
