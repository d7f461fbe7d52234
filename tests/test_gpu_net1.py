"""The forecast network LGUnet_all_1 on the B200, through the C ABI (vv_net1_*, vv_test_attn1): the SD_attn core
(rope2, rolled windows, 0 / -inf latitude mask, whole-grid stage) against the oracle's fp32 formulation, the whole network against
the reference-generated golden `net1_mid.npz` and the oracle, `integrate` against the oracle's restatement of da_4dvar.py:666-681,
and one application at the reference's real size (69 x 721 x 1440)."""
import ctypes as C
import pathlib
import sys
import time

import numpy as np
import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vaevar_b200 import _lib, build
    build.build()
    return _lib.load()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _attn_ref(qkv, gh, gw, win, shift, heads, hd):
    """SD_attn.forward between the qkv Linear and the proj Linear (Attention.py:553-641) in fp32, with the oracle's helpers."""
    from oracle.lgunet1 import _partition, _reverse, rope2, rope2_tables, shift_mask
    d = heads * hd
    x = qkv.float().view(1, gh, gw, 3 * d)
    mask = None if (shift[1] == 0 or win[1] == gw) else shift_mask(gh, gw, win, shift).to(qkv.device)
    xs = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2)) if shift[1] > 0 else x
    xw = _partition(xs, win).reshape(-1, win[0] * win[1], 3 * d)
    B_, N, _ = xw.shape
    q, k, v = xw.reshape(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0)
    tab = tuple(t.to(qkv.device) if torch.is_tensor(t) else t for t in rope2_tables(win, hd))
    q = rope2(q.reshape(-1, win[0], win[1], hd), tab).reshape(B_, heads, N, hd)
    k = rope2(k.reshape(-1, win[0], win[1], hd), tab).reshape(B_, heads, N, hd)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if mask is not None:
        attn = attn + mask.unsqueeze(1)
    out = (attn.softmax(dim=-1) @ v).transpose(1, 2).reshape(B_, N, d)
    xs = _reverse(out.reshape(-1, win[0], win[1], d), win, gh, gw)
    if shift[0] > 0:
        xs = torch.roll(xs, shifts=(shift[0], shift[1]), dims=(1, 2))
    return xs.reshape(gh * gw, d)


def _rope_table(win, hd, dev):
    """[wh * ww][hd / 2] (cos, sin) pairs in the layout of RopeArgs::table, from the oracle's tables (positional_encodings.py:231-252)."""
    from oracle.lgunet1 import rope2_tables
    sin1, cos1, sin2, cos2, d1, d2 = rope2_tables(win, hd)
    cs = torch.cat([cos1, cos2], -1).reshape(win[0] * win[1], hd // 2)
    sn = torch.cat([sin1, sin2], -1).reshape(win[0] * win[1], hd // 2)
    return torch.stack([cs, sn], -1).float().contiguous().to(dev)


@pytest.mark.parametrize("gh,gw,win,shifted,heads,hd,tc", [
    (12, 24, (6, 12), False, 3, 32, False), (12, 24, (6, 12), True, 3, 32, False), (12, 24, (6, 12), True, 2, 64, False), (12, 24, (6, 12), True, 2, 192, False),
    (6, 24, (6, 12), True, 2, 32, False),          # a single window row: every window is a masked one
    (6, 12, (6, 12), True, 1, 32, False),          # the window spans the whole width: rolled, but never masked (Attention.py:553)
    (12, 24, (12, 24), False, 2, 192, False),      # whole-grid stage, 288 tokens: the online-softmax path
    (18, 36, (18, 36), False, 1, 64, False),       # 648 tokens: key and query tiles with a ragged tail
    (30, 60, (30, 60), False, 2, 32, False),       # 1800 tokens
    (18, 36, (18, 36), False, 1, 192, True),    # the tcgen05 kernel of the whole-grid stage: 648 tokens = 2.5 query blocks, 10.1 key tiles
    (30, 60, (30, 60), False, 2, 192, True),    # 1800 tokens, two heads
    (45, 90, (45, 90), False, 6, 192, True),    # 4050 tokens, six heads (the shipped trunk width 1152)
])
def test_sd_attn_core_against_oracle(lib, gh, gw, win, shifted, heads, hd, tc):
    from vaevar_b200 import _lib
    dev = "cuda:0"
    g = torch.Generator(device="cpu").manual_seed(gh * 1000 + gw + hd + heads)
    d = heads * hd
    qkv = (torch.randn(gh * gw, 3 * d, generator=g) * 1.5).to(dev).half().contiguous()
    shift = (win[0] // 2, win[1] // 2) if shifted else (0, 0)
    ref = _attn_ref(qkv, gh, gw, win, shift, heads, hd)
    out = torch.empty(gh * gw, d, device=dev, dtype=torch.float16)
    tab = _rope_table(win, hd, dev)
    work = qkv.clone()                       # rope2 rotates q and k in place
    mask = int(shift[1] > 0 and win[1] != gw)
    _lib.check(lib.vv_test_attn1(C.c_void_p(work.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(tab.data_ptr()), gh, gw, win[0], win[1],
                                 shift[0], shift[1], heads, hd, mask, int(tc), None))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    err = _rel(out.float(), ref)
    print(f"SD_attn {gh}x{gw} win {win} shift {shift} heads {heads} hd {hd}{' [tcgen05]' if tc else ''}: rel {err:.2e}")
    assert err < 5e-3             # fp16 q / k after the rotation, fp16 P and output; fp32 softmax and accumulation
    assert torch.equal(work[:, 2 * d:], qkv[:, 2 * d:])           # v is untouched


@pytest.fixture(scope="module")
def mid_net(lib):
    from vaevar_b200.config import FORECAST_MID
    from vaevar_b200.forecast import ForecastNet
    from vaevar_b200.synth import make_state_dict_net1
    sd = make_state_dict_net1(FORECAST_MID, seed=11, rich=True)
    net = ForecastNet(FORECAST_MID, keep_out=69)
    net.load_state_dict(sd)
    net.finalize()
    return net, sd


def test_forecast_network_against_reference_golden(mid_net, gold):
    """net1_mid.npz holds outputs of the reference's own LGUnet_all_1 (tools/make_golden.py::golden_lgunet1_mid)."""
    from oracle.lgunet1 import lgunet1_forward
    from vaevar_b200.config import FORECAST_MID
    net, sd = mid_net
    g = gold("net1_mid.npz")
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(int(g["x_seed"]))).standard_normal((1, 69, *FORECAST_MID.img_size), dtype=np.float32))
    y = net.forward(x[0].cuda())
    torch.cuda.synchronize()
    assert tuple(y.shape) == (69, 97, 192) and torch.isfinite(y).all()
    got = y.flatten()[torch.from_numpy(g["y_idx"]).cuda()].cpu().double().numpy()
    want = g["y_val"].astype(np.float64)
    err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    print(f"LGUnet_all_1 mid forward vs reference golden: rel {err:.2e}; launches {net.last_launch_count}; |y|_1 rel "
          f"{abs(float(y.double().abs().sum()) / float(g['y_abs']) - 1):.2e}")
    assert err < 1e-2
    assert abs(float(y.double().abs().sum()) / float(g["y_abs"]) - 1) < 2e-3
    # every element, per channel, against the oracle (pinned to the same golden on CPU)
    with torch.no_grad():
        yr = lgunet1_forward(x, {k: torch.from_numpy(v) for k, v in sd.items()}, FORECAST_MID)[0, :69]
    per_chan = ((y.cpu().double() - yr.double()) ** 2).sum((1, 2)).sqrt() / (yr.double() ** 2).sum((1, 2)).sqrt()
    print(f"  per-channel rel error: max {float(per_chan.max()):.2e} (channel {int(per_chan.argmax())})")
    assert float(per_chan.max()) < 2e-2


def test_forecast_network_all_outputs_and_module_shell(lib, mid_net):
    """keep_out = 0 produces the mean | std halves of every group (138 channels); the nn.Module shell has the reference's call surface
    (constructor keywords, state_dict names, batched forward without autograd)."""
    from oracle.lgunet1 import lgunet1_forward
    from vaevar_b200.config import FORECAST_MID
    from vaevar_b200.modules import LGUnet_all_1
    _, sd = mid_net
    m = LGUnet_all_1(**FORECAST_MID.to_reference_kwargs())
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not missing and not unexpected
    m = m.cuda().eval()
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(5)).standard_normal((2, 69, *FORECAST_MID.img_size), dtype=np.float32))
    y = m(x.cuda())
    assert tuple(y.shape) == (2, 138, 97, 192) and not y.requires_grad
    with torch.no_grad():
        yr = lgunet1_forward(x, {k: torch.from_numpy(v) for k, v in sd.items()}, FORECAST_MID)
    err = _rel(y.cpu(), yr)
    print(f"LGUnet_all_1 shell, batch 2, 138 channels: rel {err:.2e}")
    assert err < 1e-2


def test_forecast_integrate_against_oracle(mid_net):
    """cyclic_4dvar.integrate(xa, forecast_model, 2) (da_4dvar.py:666-681): normalise, two applications keeping [:69], de-normalise."""
    from oracle.lgunet1 import lgunet1_forward
    from vaevar_b200.config import FORECAST_MID, era5_stats
    net, sd = mid_net
    mean, std, _ = era5_stats()
    mean_t = torch.from_numpy(mean).float().reshape(69, 1, 1)
    std_t = torch.from_numpy(std).float().reshape(69, 1, 1)
    rng = np.random.Generator(np.random.PCG64(9))
    xa = mean_t + std_t * torch.from_numpy(rng.standard_normal((69, *FORECAST_MID.img_size), dtype=np.float32))
    out = net.integrate(xa.cuda(), 2)
    torch.cuda.synchronize()
    sdt = {k: torch.from_numpy(v) for k, v in sd.items()}
    with torch.no_grad():
        z = ((xa - mean_t) / std_t).unsqueeze(0)
        for _ in range(2):
            z = lgunet1_forward(z, sdt, FORECAST_MID)[:, :69]
        ref = z[0] * std_t + mean_t
    dn = (out.cpu() - mean_t) / std_t                     # compare in normalised units (channel magnitudes differ by 1e6)
    rn = (ref - mean_t) / std_t
    err = _rel(dn, rn)
    print(f"integrate(xa, LGUnet_all_1, 2): rel {err:.2e}")
    assert err < 2e-2


def test_forecast_network_at_the_reference_size(lib):
    """One application of the shipped 0.25-degree configuration (69 x 721 x 1440; 259 200 / 64 800 / 16 200 tokens, a 16 200-token
    whole-grid stage): finite output, and the same answer from two runs (the plan holds no state between applications)."""
    from vaevar_b200.config import FORECAST_FULL
    from vaevar_b200.forecast import ForecastNet
    from vaevar_b200.synth import make_state_dict_net1
    t0 = time.time()
    sd = make_state_dict_net1(FORECAST_FULL, seed=1, rich=False)
    net = ForecastNet(FORECAST_FULL, keep_out=69)
    net.load_state_dict(sd)
    net.finalize()
    del sd
    x = torch.randn(69, 721, 1440, generator=torch.Generator().manual_seed(3)).cuda()
    y0 = net.forward(x)
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    net.forward(x)                      # the timed application is the third: allocator and clocks have settled
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    y1 = net.forward(x)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"LGUnet_all_1 at 721x1440: {ms:.1f} ms per application, {net.last_launch_count} launches, {net.device_bytes / 2**30:.2f} GiB on the device "
          f"(setup {t_setup:.0f} s); 16.1 TFLOP -> {16.1 / ms * 1e3:.0f} TFLOP/s")
    assert tuple(y1.shape) == (69, 721, 1440) and torch.isfinite(y1).all()
    assert torch.equal(y0, y1)
    assert float(y1.std()) > 1e-3
    net.close()
