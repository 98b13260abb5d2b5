"""SURVEY 8f item 3, host logic of the callers: `patch_model` (helper.py:34-85) walks a model, names its modules, hands
every torch.nn.Linear whose qualified name contains none of `skip_modules` to the processor and leaves the others alone.
A recording stand-in processor keeps this on the CPU; what the real processors build is covered by the GPU suite."""
import torch

from gemlite_b200 import helper


class _Packed(torch.nn.Module):
    def __init__(self, src):
        super().__init__()
        self.shape = tuple(src.weight.shape)
        self.src_name = src.name


class _Recorder:
    def __init__(self):
        self.seen = []

    def from_linear(self, layer):
        self.seen.append((layer.name, tuple(layer.weight.shape)))
        return _Packed(layer)


class _Block(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.q_proj = torch.nn.Linear(32, 32, bias=False)
        self.mlp = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.GELU(), torch.nn.Linear(64, 32))
        self.norm = torch.nn.LayerNorm(32)


class _Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList([_Block(), _Block()])
        self.vision_tower = torch.nn.Sequential(torch.nn.Linear(32, 32))
        self.lm_head = torch.nn.Linear(32, 100, bias=False)


def test_patch_model_replaces_linears_and_respects_skip_list():
    rec = _Recorder()
    model = helper.patch_model(_Model(), device="cpu", processor=rec)
    want = [(f"layers.{i}.{n}", s) for i in range(2) for n, s in
            (("q_proj", (32, 32)), ("mlp.0", (64, 32)), ("mlp.2", (32, 64)))]
    assert rec.seen == want
    for i in range(2):
        blk = model.layers[i]
        assert isinstance(blk.q_proj, _Packed) and isinstance(blk.mlp[0], _Packed) and isinstance(blk.mlp[2], _Packed)
        assert isinstance(blk.mlp[1], torch.nn.GELU) and isinstance(blk.norm, torch.nn.LayerNorm)
        assert blk.q_proj.src_name == f"layers.{i}.q_proj"
    # default skip_modules = ["lm_head", "vision", "visual"]: untouched torch.nn.Linear modules
    assert type(model.lm_head) is torch.nn.Linear and type(model.vision_tower[0]) is torch.nn.Linear
    # every module got its qualified name (the reference tags them the same way, helper.py:49-50)
    assert model.lm_head.name == "lm_head" and model.layers[1].mlp.name == "layers.1.mlp"

    rec2 = _Recorder()
    helper.patch_model(_Model(), device="cpu", processor=rec2, skip_modules=["mlp"])
    assert [n for n, _ in rec2.seen] == ["layers.0.q_proj", "layers.1.q_proj", "vision_tower.0", "lm_head"]


def test_default_batch_sizes_and_processor_surface():
    assert helper.default_batch_sizes == [4096, 2048, 1024, 512, 256, 128, 64, 32, 16, 8, 4, 2, 1]     # helper.py:1065
    for name in ("A16W8", "A16W8_INT8", "A16W8_FP8", "A16Wn", "A16W8_HQQ_INT", "A16W4_HQQ_INT", "A16W2_HQQ_INT",
                 "A16W1_HQQ_INT", "A16W4_MXFP", "A16W8_MXFP", "A8W8_dynamic", "A8W8_int8_dynamic", "A8W8_fp8_dynamic",
                 "A8W8_INT8_dynamic", "A8W8_FP8_dynamic", "A8Wn_HQQ_INT_dynamic", "A8W4_HQQ_INT_dynamic",
                 "A8W2_HQQ_INT_dynamic", "A8W8_MXFP_dynamic", "A8Wn_MXFP_dynamic", "A8W4_MXFP_dynamic",
                 "A4W4_MXFP_dynamic", "A4W4_NVFP_dynamic", "A16W158_INT", "A8W158_INT_dynamic", "patch_model", "warmup"):
        assert hasattr(helper, name), name


def test_forward_functional_traces_with_fake_tensors():
    """The registered fake of `forward_functional` (core.py:197-206 in the reference) lets torch.compile / vLLM trace a
    packed layer without running it: shape (..., N), dtype = the layer's OUTPUT dtype (the reference's fake returns
    x.dtype, which is wrong for int8 / fp8 activations -- not reproduced)."""
    import os
    from torch._subclasses.fake_tensor import FakeTensorMode
    import gemlite_b200  # noqa: F401  (registers the op)
    from gemlite_b200 import DType
    ns = os.environ.get("GEMLITE_B200_OP_NAMESPACE", "gemlite")
    op = getattr(torch.ops, ns).forward_functional
    K, N = 256, 64
    with FakeTensorMode():
        W_q = torch.empty(K // 8, N, dtype=torch.int32)
        scales = torch.empty(K // 64, N, dtype=torch.float16)
        zeros = torch.empty(K // 64, N, dtype=torch.float16)
        meta = [0, 4, 64, 15, 8, DType.FP16.value, DType.FP16.value, DType.FP32.value, DType.FP16.value, 0, 4, 1]
        y = op(torch.empty(3, 5, K, dtype=torch.float16), None, [W_q, scales, zeros], meta, -1)
        assert tuple(y.shape) == (3, 5, N) and y.dtype == torch.float16
        meta8 = [1, 8, K, 255, 1, DType.INT8.value, DType.BF16.value, DType.INT32.value, DType.FP32.value, 3, 0, 0]
        y8 = op(torch.empty(7, K, dtype=torch.bfloat16), torch.empty(N, dtype=torch.bfloat16),
                [torch.empty(K, N, dtype=torch.int8), torch.empty(1, N, dtype=torch.float32), torch.empty(1, 0, dtype=torch.int32)], meta8, 4)
        assert tuple(y8.shape) == (7, N) and y8.dtype == torch.bfloat16
