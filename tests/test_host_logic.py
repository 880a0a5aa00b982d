"""Host-side logic on CPU: sharding, report aggregation / formatting, and the N>1 path of evaluate()
with world_size-2 gloo processes (the per-batch numbers come from the oracle here; on GPUs they come
from the fused metrics kernel - tests/test_gpu_model.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from imageenhancement_mp_b200 import data_utils as du
from imageenhancement_mp_b200 import dist as idist
from imageenhancement_mp_b200 import eval as ieval
from imageenhancement_mp_b200 import synth

T = 4


def oracle_step_totals(model, xb, xt, burst_length):
    """CPU stand-in for eval.gpu_step_totals: same additive totals layout, numbers from the oracle."""
    recon = model(xb)
    n = xb.shape[0]
    tot = torch.zeros(burst_length + 6, dtype=torch.float64)
    for i in range(n):
        s = oracle.eval_step(recon[i:i + 1], xb[i:i + 1], xt[i:i + 1], burst_length)
        tot[0] += s["psnr"]
        for t in range(burst_length):
            tot[1 + t] += s["psnr_perlayer"][t]
        tot[burst_length + 1] += s["psnr_noise0"]
        tot[burst_length + 2] += s["psnr_average"]
        tot[burst_length + 3] += s["loss1"]
        tot[burst_length + 4] += s["perlayer_loss"]
        tot[burst_length + 5] += 1
    return tot


def fake_model(xb):
    """Deterministic stand-in network: frame average + per-frame copies (shape [N,H,W,T+1])."""
    burst = xb[..., :T]
    return torch.cat([burst.mean(-1, keepdim=True), burst], dim=-1)


def make_batches(nb, n, h, w):
    return [synth.make_batch(n, h, w, seed=100 + i) for i in range(nb)]


def reference_report(batches):
    steps = [oracle.eval_step(fake_model(x), x, t, T) for x, t in batches]
    return oracle.eval_report(steps, T)


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 8, 256):
        for world in (1, 2, 3, 8):
            spans = [idist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_spatial_shards_tile_the_image():
    for H in (2448, 720, 256, 64):
        for world in (1, 2, 3, 8):
            sh = idist.spatial_shards(H, world)
            assert sh[0]["own"][0] == 0 and sh[-1]["own"][1] == H
            for i, d in enumerate(sh):
                a, b = d["own"]; s0, s1 = d["slab"]
                assert a % 8 == 0 and b % 8 == 0 and s0 % 8 == 0 and s1 % 8 == 0 and s0 <= a <= b <= s1
                assert s0 == max(0, a - idist.SPATIAL_HALO) and s1 == min(H, b + idist.SPATIAL_HALO)
                assert d["own_in_slab"] == (a - s0, b - s0)
                if i:
                    assert sh[i - 1]["own"][1] == a


def test_single_process_report_matches_oracle_aggregation():
    batches = make_batches(3, 4, 40, 48)
    params = dict(synth.DEFAULT_PARAMS)
    lines = []
    rep = ieval.evaluate(fake_model, batches, params, step=7, out=lines.append, step_totals=oracle_step_totals)
    ref = reference_report(batches)
    for k in ieval.REPORT_KEYS:
        assert rep[k] == pytest.approx(ref[k], rel=1e-6, abs=1e-6), k
    assert rep["val_psnrnoshow0_unbiased"] == pytest.approx(ref["val_psnrnoshow0_unbiased"], rel=1e-6)
    assert len(lines) == 8 and lines[1].startswith("epoch 7: val_deblur_loss = ")
    assert [l.split(":")[1].split("=")[0].strip() for l in lines[1:]] == list(ieval.REPORT_KEYS)


def test_ps_lines_of_the_report():
    """params["ps"] (eval.py:103,159-162,189-191): the step returns (totals, sum over images of cost_volume); the report
    gains `variance loss` / `variance` after val_total_loss, like the reference prints them."""
    batches = make_batches(3, 4, 40, 48)
    params = dict(synth.DEFAULT_PARAMS, ps=True)
    g = torch.Generator().manual_seed(3)
    bases = [torch.softmax(torch.randn(4, 15 * 15 * T, 10, generator=g, dtype=torch.float64) * 5, dim=1)
             .view(4, 15, 15, T, 10) for _ in batches]
    it = iter(bases)

    def step(model, xb, xt, burst_length):
        bas = next(it)
        cv = sum(oracle.cost_volume(bas[i:i + 1]) for i in range(bas.shape[0]))
        return oracle_step_totals(model, xb, xt, burst_length), cv.reshape(1)

    lines = []
    rep = ieval.evaluate(fake_model, batches, params, step=2, out=lines.append, step_totals=step, beta_coef=50.0)
    ref = float(torch.stack([oracle.cost_volume(b) for b in bases]).mean())     # Keras Mean of the batch values
    assert rep["variance"] == pytest.approx(ref, rel=1e-9)
    assert rep["variance loss"] == pytest.approx(50.0 * ref, rel=1e-9)
    names = [l.split(":")[1].split("=")[0].strip() for l in lines[1:]]
    assert names == list(ieval.REPORT_KEYS[:3]) + ["variance loss", "variance"] + list(ieval.REPORT_KEYS[3:])
    base = reference_report(batches)
    for k in ieval.REPORT_KEYS:
        assert rep[k] == pytest.approx(base[k], rel=1e-6, abs=1e-6), k


def test_totals_layout_roundtrip():
    tot = torch.arange(T + 6, dtype=torch.float64) + 1
    tot[-1] = 2
    r = du.totals_to_report(tot, T)
    assert r["psnr"] == 0.5 and r["psnr_perlayer"] == [1.0, 1.5, 2.0, 2.5] and r["count"] == 2
    assert r["psnr_noise0"] == 3.0 and r["psnr_average"] == 3.5 and r["loss1"] == 4.0 and r["perlayer_loss"] == 4.5


def test_totals_layout_with_ssim_extension():
    """T+7 totals (ie_metric_totals_ssim_f64): the mean-SSIM sum sits before the image count; the reference's numbers
    keep their places and `val_ssim` appears only then."""
    tot = torch.arange(T + 7, dtype=torch.float64) + 1
    tot[-1] = 2
    r = du.totals_to_report(tot, T)
    assert r["psnr"] == 0.5 and r["perlayer_loss"] == 4.5 and r["count"] == 2
    assert r["ssim"] == (T + 6) / 2
    rep = ieval.make_report(tot, 3, T)
    assert rep["val_ssim"] == (T + 6) / 2 and rep["val_psnr"] == 0.5
    assert "val_ssim" not in ieval.make_report(torch.cat([tot[:T + 5], tot[-1:]]), 3, T)
    with pytest.raises(AssertionError):
        du.totals_to_report(torch.ones(T + 8, dtype=torch.float64), T)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, w, _ = idist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    batches = make_batches(3, 5, 40, 48)          # 5 images per batch: uneven 3/2 split across two ranks
    rep = ieval.evaluate(fake_model, batches, dict(synth.DEFAULT_PARAMS), out=None, step_totals=oracle_step_totals)
    q.put((rank, {k: rep[k] for k in ieval.REPORT_KEYS + ("count",)}))
    dist.barrier()
    dist.destroy_process_group()


def _ps_bases(nb, n):
    g = torch.Generator().manual_seed(17)
    return [torch.softmax(torch.randn(n, 15 * 15 * T, 10, generator=g, dtype=torch.float64) * 5, dim=1)
            .view(n, 15, 15, T, 10) for _ in range(nb)]


def _worker_ps(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    idist.init_from_env(backend="gloo")
    batches = make_batches(2, 5, 40, 48)
    bases = iter(_ps_bases(2, 5))

    def step(model, xb, xt, burst_length):
        lo, hi = idist.shard_range(5, rank, world)           # this rank's images of the batch, like shard_batch
        bas = next(bases)[lo:hi]
        cv = sum(oracle.cost_volume(bas[i:i + 1]) for i in range(bas.shape[0]))
        return oracle_step_totals(model, xb, xt, burst_length), cv.reshape(1)

    rep = ieval.evaluate(fake_model, batches, dict(synth.DEFAULT_PARAMS, ps=True), out=None, step_totals=step)
    q.put((rank, {k: rep[k] for k in ("variance", "variance loss", "count", "val_psnr")}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_ps_variance_rides_in_the_all_reduce():
    """params["ps"] under two ranks with an uneven 3/2 split: the per-rank cost_volume sums are appended to the totals
    vector, all-reduced once, and `variance` equals the mean over ALL images on both ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ps, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    per_image = [oracle.cost_volume(b[i:i + 1]) for b in _ps_bases(2, 5) for i in range(5)]
    ref = float(torch.stack(per_image).mean())
    base = reference_report(make_batches(2, 5, 40, 48))
    for rank in (0, 1):
        assert got[rank]["count"] == 10
        assert got[rank]["variance"] == pytest.approx(ref, rel=1e-9)
        assert got[rank]["variance loss"] == pytest.approx(100.0 * ref, rel=1e-9)
        assert got[rank]["val_psnr"] == pytest.approx(base["val_psnr"], rel=1e-6)


def _worker_batch1(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    idist.init_from_env(backend="gloo")
    batches = make_batches(3, 1, 40, 48)          # the reference's default batch_size = 1 (eval.py:37): rank 1 gets nothing
    rep = ieval.evaluate(fake_model, batches, dict(synth.DEFAULT_PARAMS), out=None, step_totals=oracle_step_totals)
    q.put((rank, {k: rep[k] for k in ieval.REPORT_KEYS + ("count",)}))
    # no data on ANY rank: every rank raises, after the collective
    try:
        ieval.evaluate(fake_model, [], dict(synth.DEFAULT_PARAMS), out=None, step_totals=oracle_step_totals)
        q.put((rank + 10, "no error"))
    except ValueError as e:
        q.put((rank + 10, str(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_batch_of_one_does_not_hang():
    """with batch_size 1 and two ranks, rank 1's shard of every batch is empty.  It must still enter the
    all-reduce (with zeros) instead of raising before it, and both ranks report the single-process numbers."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_batch1, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(4))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = reference_report(make_batches(3, 1, 40, 48))
    for rank in (0, 1):
        assert got[rank]["count"] == 3
        for k in ieval.REPORT_KEYS:
            assert got[rank][k] == pytest.approx(ref[k], rel=1e-6, abs=1e-6), (rank, k)
        assert "no validation data on any rank" in got[rank + 10]


def test_two_rank_gloo_evaluate_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = reference_report(make_batches(3, 5, 40, 48))
    for rank in (0, 1):
        assert got[rank]["count"] == 15
        for k in ieval.REPORT_KEYS:
            assert got[rank][k] == pytest.approx(ref[k], rel=1e-6, abs=1e-6), (rank, k)


def test_draw_burst_params_follows_the_reference_crops():
    """draw_burst_params: ranges and structure of the reference's random crops (data_utils.py:432-457), the absolute
    origins against the oracle's frame_origins for the same draws, determinism, and the symmetric zero padding of a
    source smaller than the crop (negative origins)."""
    from oracle import preprocess as opre
    params = dict(synth.DEFAULT_PARAMS, height=24, width=32, BURST_LENGTH=4)
    up, jit, sj = params["upscale"], params["jitter"], params["smalljitter"]
    g = torch.Generator().manual_seed(3)
    n, hs, ws = 64, 400, 500
    d = du.draw_burst_params(n, (hs, ws), params, generator=g)
    h_up, w_up = 24 * up + 2 * jit * up, 32 * up + 2 * jit * up
    assert d["org"].dtype == torch.int32 and d["org"].shape == (n, 4, 2)
    assert int(d["crop0"][:, 0].min()) >= 0 and int(d["crop0"][:, 0].max()) <= hs - h_up
    assert int(d["crop0"][:, 1].min()) >= 0 and int(d["crop0"][:, 1].max()) <= ws - w_up
    big, off = d["use_big"], d["frame_off"]
    assert int(off[big].max()) <= 2 * jit * up and int(off[~big].max()) <= 2 * sj * up and int(off.min()) >= 0
    assert bool(big.any()) and bool((~big).any())
    for i in range(n):
        draws = {"crop0": tuple(int(v) for v in d["crop0"][i]), "use_big": [bool(v) for v in big[i]],
                 "frame_off": [tuple(int(v) for v in o) for o in off[i]]}
        assert opre.frame_origins(params, draws) == [tuple(int(v) for v in o) for o in d["org"][i]]
    # every frame's (h*up x w*up) window stays inside the first crop, hence inside the source
    assert int(d["org"].min()) >= 0
    assert int(d["org"][..., 0].max()) + 24 * up <= hs and int(d["org"][..., 1].max()) + 32 * up <= ws
    for k, lo, hi in (("white_level", 0.1, 1.0), ("sig_read", 10 ** -3, 10 ** -1.5), ("sig_shot", 10 ** -2, 10 ** -1)):
        assert d[k].shape == (n,) and float(d[k].min()) >= lo * (1 - 1e-6) and float(d[k].max()) <= hi * (1 + 1e-6)
    d2 = du.draw_burst_params(n, (hs, ws), params, generator=torch.Generator().manual_seed(3))
    assert all(torch.equal(d[k], d2[k]) for k in d)
    # source smaller than the crop: the reference pads (h_up - Hs + 1)//2 on both sides (:435-438) -> fixed negative origin
    small = du.draw_burst_params(4, (100, 120), params, generator=g)
    assert torch.equal(small["crop0"], torch.tensor([[-((h_up - 100 + 1) // 2), -((w_up - 120 + 1) // 2)]] * 4))


# ---------------------------------------------------------------------------------------------------------------------
# Invariants the CUDA kernels rely on, restated in Python (the device code cannot run here; these pin the arithmetic
# the host side and the kernels share).

def _fastdiv_magic(d):
    """make_fastdiv of csrc/conv_tcgen05.cu: q = umulhi(n, m) >> s for 0 <= n < 2^31, d >= 2."""
    l = 0
    while (1 << l) < d:
        l += 1
    m = ((1 << (31 + l)) // d) + 1
    assert m < (1 << 32), d
    return m, l - 1


def test_multiply_high_division_is_exact():
    """The epilogues / builders / stagers decode raster rows with a multiply-high instead of a division
    (EpiParams::plane_m ...): exact for every dividend below 2^31 and every divisor the rasters produce."""
    import random
    rnd = random.Random(7)
    divisors = list(range(2, 700)) + [(h + 1) * (w + 1) for h, w in ((8, 8), (104, 104), (720, 1280), (2448, 3264))] \
        + [rnd.randrange(2, 1 << 30) for _ in range(500)] + [1 << k for k in range(1, 31)] + [(1 << k) + 1 for k in range(1, 30)]
    for d in divisors:
        m, s = _fastdiv_magic(d)
        for n in (0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 31) - d, rnd.randrange(1 << 31), rnd.randrange(1 << 31)):
            if 0 <= n < (1 << 31):
                assert ((n * m) >> 32) >> s == n // d, (n, d)


@pytest.mark.parametrize("n,h,w,hs,ws,c", [(3, 104, 104, 100, 100, 5), (2, 32, 32, 32, 32, 5), (5, 8, 8, 8, 8, 5), (4, 8, 8, 5, 6, 3),
                                           (2, 16, 24, 13, 20, 10), (1, 40, 300, 40, 300, 5), (2, 40, 200, 33, 190, 3),
                                           (4, 4, 4, 2, 2, 5), (8, 2, 2, 1, 1, 3)])
def test_first_layer_staging_plan_covers_every_read(n, h, w, hs, ws, c):
    """conv_first_staged_kernel (csrc/conv_tcgen05.cu): per 128-row tile and filter row the stager copies ONE contiguous
    run of source pixels [lo - 1, hi + 1] (16-byte aligned) into a slot of 130 pixels; the builders then read pixel
    (y + dy - 1, x + dx - 1) at a slot-relative offset.  Restated with numpy: the run never exceeds the slot and every
    read a builder makes lands inside its run at the right element."""
    import numpy as np
    wp, plane = w + 1, (h + 1) * (w + 1)
    R = n * plane
    x = np.arange(n * hs * ws * c, dtype=np.int64)                      # element index = its own value
    total = x.size
    if total % 4:
        pytest.skip("the staged kernel needs a float count that is a multiple of 4 (the gather kernel runs instead)")
    slot_bytes = ((130 * c * 4 + 32 + 127) // 128) * 128
    r = np.arange(((R + 127) // 128) * 128)
    img, pr = r // plane, r % plane
    yp, xx = pr // wp, pr % wp
    y = yp - 1
    interior = (r < R) & (y >= 0) & (y < h) & (xx < w)
    for dy in range(3):
        sy = y + dy - 1
        rowok = interior & (sy >= 0) & (sy < hs)
        v = (img * hs + sy) * ws + np.minimum(xx, ws - 1)
        vt = np.where(rowok, v, -1).reshape(-1, 128)
        hi = vt.max(axis=1)
        lo = np.where(vt < 0, 1 << 60, vt).min(axis=1)
        for t in range(vt.shape[0]):
            if hi[t] < 0:
                continue                                                  # no reader in this tile: no copy
            f0, f1 = max((lo[t] - 1) * c, 0), min((hi[t] + 2) * c, total)
            b0 = (f0 * 4) & ~15
            nbytes = ((f1 * 4 + 15) & ~15) - b0
            assert nbytes <= slot_bytes, (t, dy, nbytes, slot_bytes)
            assert b0 + nbytes <= total * 4
            run = x[b0 // 4:(b0 + nbytes) // 4]
            rows = np.nonzero(rowok.reshape(-1, 128)[t])[0] + t * 128
            for g in range(3):
                sx = xx[rows] + g - 1
                ok = (sx >= 0) & (sx < ws)
                off = ((img[rows] * hs + sy[rows]) * ws + (xx[rows] - 1)) * c - b0 // 4 + g * c
                off, want = off[ok], ((img[rows] * hs + sy[rows]) * ws + sx)[ok] * c
                assert (off >= 0).all() and (off + c <= run.size).all()
                assert (run[off] == want).all()
