"""SURVEY 8f item 3 on the GPU (only exercised on CPU with stubs): `patch_model` with real processors on a
small transformer-ish model, `warmup`, and a `torch.compile`d packed layer / patched model -- every Linear the reference
would replace (helper.py:34-85) becomes a packed GemLiteLinear running the sm_100a kernels, outputs stay within the
quantisation error of the float model, the compiled graph goes through the registered custom op (core.py:128-206)
without a graph break and reproduces the eager result."""
import pytest
import torch

import gemlite_b200
from gemlite_b200 import GemLiteLinear, _cabi, helper

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Block(torch.nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.q_proj = torch.nn.Linear(d, d, bias=False)
        self.up = torch.nn.Linear(d, h, bias=True)
        self.down = torch.nn.Linear(h, d, bias=True)
        self.norm = torch.nn.LayerNorm(d)

    def forward(self, x):
        x = x + self.q_proj(self.norm(x))
        return x + self.down(torch.nn.functional.gelu(self.up(x)))


class _Model(torch.nn.Module):
    def __init__(self, d=1024, h=2048, n=2):
        super().__init__()
        self.layers = torch.nn.ModuleList([_Block(d, h) for _ in range(n)])
        self.lm_head = torch.nn.Linear(d, 512, bias=False)

    def forward(self, x):
        for blk in self.layers:
            x = blk(x)
        return self.lm_head(x)


def _model(dtype):
    torch.manual_seed(0)
    m = _Model().to(dtype)
    for p in m.parameters():
        p.requires_grad = False
    return m


@pytest.mark.parametrize("proc_name,tol", [("A16W8_INT8", 2e-2), ("A8W8_INT8_dynamic", 4e-2), ("A16W4_MXFP", 2.5e-1), ("A8W8_FP8_dynamic", 8e-2)])
def test_patch_model_runs_every_linear_on_the_cuda_path(proc_name, tol):
    dtype = torch.bfloat16 if "MXFP" in proc_name else torch.float16
    ref = _model(dtype).to(DEV)
    x = (torch.randn(3, 7, 1024, generator=torch.Generator().manual_seed(1)) / 4).to(device=DEV, dtype=dtype)
    y_ref = ref(x).float()
    model = helper.patch_model(_model(dtype), device=DEV, processor=getattr(helper, proc_name)(device=DEV, dtype=dtype))
    packed = [m for m in model.modules() if isinstance(m, GemLiteLinear)]
    assert len(packed) == 6 and type(model.lm_head) is torch.nn.Linear          # lm_head is in the default skip list
    assert all(p.W_q.is_cuda for p in packed)
    n0 = _cabi.launch_count()
    y = model(x).float()
    assert _cabi.launch_count() - n0 >= 6                                         # our kernels ran, not torch.nn.Linear
    rel = ((y - y_ref).norm() / y_ref.norm()).item()
    assert rel < tol, (proc_name, rel)
    # decode-sized call through the same patched model
    y1 = model(x[:1, :1])
    assert y1.shape == (1, 1, 512) and torch.isfinite(y1.float()).all()


def test_warmup_runs_every_batch_size():
    n0 = _cabi.launch_count()
    helper.warmup(helper.A16W8_INT8(device=DEV, dtype=torch.float16), shapes=[(512, 1024)], batch_sizes=[64, 16, 1])
    helper.warmup(helper.A8W8_INT8_dynamic(device=DEV, dtype=torch.float16), shapes=[(512, 1024)], batch_sizes=[200, 1])
    assert _cabi.launch_count() - n0 >= 5


def test_torch_compile_goes_through_the_registered_op():
    """fullgraph=True: a graph break at the packed layer would raise.  The compiled module must reproduce eager."""
    dtype = torch.float16
    g = torch.Generator().manual_seed(3)
    W_q = torch.randint(0, 16, (512, 1024), generator=g, dtype=torch.int32).to(torch.uint8)
    s = (torch.rand(512 * 1024 // 128, 1, generator=g) * 0.01 + 0.001).to(dtype)
    z = (torch.rand(512 * 1024 // 128, 1, generator=g) * 15).to(dtype)
    layer = helper.A16W4_HQQ_INT(device=DEV, dtype=dtype).from_weights(W_q.to(DEV), s.to(DEV), z.to(DEV))

    class Wrap(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = layer

        def forward(self, x):
            return torch.nn.functional.silu(self.lin(x)) * 2.0

    m = Wrap()
    cm = torch.compile(m, fullgraph=True)
    for M in (1, 40, 300):
        x = (torch.randn(M, 1024, generator=g) / 10).to(device=DEV, dtype=dtype)
        y_eager = m(x)
        n0 = _cabi.launch_count()
        y_c = cm(x)
        assert _cabi.launch_count() > n0                                          # the CUDA kernel ran inside the compiled graph
        assert torch.allclose(y_c.float(), y_eager.float(), rtol=2e-3, atol=2e-3), M
