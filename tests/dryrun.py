"""CPU dry run of the plugin's launch sequence: every C-ABI wrapper of toc3d_b200.lib is replaced by a recorder and
torch.cuda's streams / events by inert stand-ins, so `_forward_core` walks through the host logic (workspaces, static
maps, per-stage tables, option switches) without a GPU.  Test infrastructure only - nothing is computed."""
import contextlib
from unittest import mock

import torch

from toc3d_b200 import backbone as BB
from toc3d_b200 import lib as L



class _Stream:
    def __init__(self, *a, **k):
        pass

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass


def _describe(v):
    if torch.is_tensor(v):
        return ("T", str(v.dtype).replace("torch.", ""), tuple(v.shape), v.data_ptr())
    if isinstance(v, dict):
        return {k: _describe(x) for k, x in sorted(v.items())}
    if isinstance(v, (list, tuple)):
        return [_describe(x) for x in v]
    return v


@contextlib.contextmanager
def recording():
    """Patches toc3d_b200.lib + torch.cuda inside the block; yields the list of (name, args, kwargs) records."""
    calls = []
    wrappers = [n for n in dir(L) if not n.startswith("_") and callable(getattr(L, n)) and not isinstance(getattr(L, n), type)
                and getattr(getattr(L, n), "__module__", None) == L.__name__]

    def recorder(name):
        def f(*a, **k):
            calls.append((name, _describe(a), _describe(k), (a, k)))
            return k.get("out")
        return f

    with contextlib.ExitStack() as st:
        for n in wrappers:
            if n == "load":
                st.enter_context(mock.patch.object(L, n, lambda: None))
            elif n == "motion_blob_floats":          # host-only size helper of the C-ABI (include/toc3d_b200.h layout)
                st.enter_context(mock.patch.object(L, n, lambda Q, C: 392 + 384 * 256 + 256 + 2 * 256 + 6 * (256 * 256 + 256)
                                                   + 2 * (180 * 256 + 256) + 256 * C + 256 + 2 * Q + 4))
            else:
                st.enter_context(mock.patch.object(L, n, recorder(n)))
        st.enter_context(mock.patch.object(torch.cuda, "Stream", _Stream))
        st.enter_context(mock.patch.object(torch.cuda, "Event", _Event))
        st.enter_context(mock.patch.object(torch.cuda, "current_stream", lambda *a, **k: _Stream()))
        st.enter_context(mock.patch.object(torch.cuda, "stream", lambda s: contextlib.nullcontext()))
        yield calls


def run(model, inputs, **options):
    """One eager forward of `model` (CPU tensors) under recording(); options are set as model attributes
    (view_groups, ...).  -> list of records."""
    for k, v in options.items():
        setattr(model, k, v)
    model.refresh_weights()
    x = inputs["x"]
    with recording() as calls, torch.no_grad():
        eng = BB._Engine(model, torch.device("cpu"))
        eng.has_cls = model.pretrain_use_cls_token
        xx, V, Hi, Wi = model._prep_img(x)
        grid = (Hi // model.patch_size, Wi // model.patch_size)
        if hasattr(model, "pruning_loc"):
            q_kw = None
            if inputs.get("prev_exists"):
                q_kw = {k: inputs[k] for k in ("temp_queries", "temp_ref_points", "temp_vel", "temp_timestamp",
                                               "temp_ego_pose", "ego_pose_inv")}
            model._forward_core(eng, xx, grid, q_kw, None, None, None)
        else:
            model._dense_core(eng, xx, grid) if hasattr(model, "_dense_core") else _dense(model, eng, xx, grid)
    return calls


def _dense(model, eng, x, grid):
    wsp = eng.workspace(x.shape[0], grid[0], grid[1])
    X = eng.stem(x, wsp)
    for i in range(len(model.blocks)):
        eng.dense_block(i, X, wsp)
    return X


def names(calls):
    out = []
    for name, a, k, _ in calls:
        if name == "gemm":
            out.append("gemm:%d" % a[2])
        else:
            out.append(name)
    return out
