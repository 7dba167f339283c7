"""GPU, world_size 2 (and 8 when present) over NCCL: the data-parallel runner on hardware.

What must hold for any number of ranks:
  * the gathered vectors (J, exceed counters; nonconformity scores, weights) are the per-rank results in global index order;
  * every rank derives the SAME metrics dictionary and the SAME quantile, bit for bit, and they equal what one process computes
    from the gathered vectors (host sort for the quantile, reference formulae for the metrics);
  * the in-kernel Philox noise is keyed by (seed, global sample index, t): the samples a rank draws for its shard agree with the
    single-rank chain over the whole batch (to 1e-4: GroupNorm statistics are reduced with atomics whose order depends on
    where a sample sits in a tile; nothing else couples samples).
Skipped when fewer GPUs are visible (the driver's single-GPU box); profiles/r02_nccl_ranks.log holds the 2- and 8-GPU runs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Cfg:
    use_max_safety = True
    u_bound = 0.8
    guidance_weights = {"w_score": 500.0}
    nt = 11
    InfFT_Q = None


def _model():
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
    return s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=4, ddim_sampling_eta=1.0, temporal=True,
                               use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()


def _inputs(n):
    from oracle import fixtures as fx
    g = torch.Generator().manual_seed(5)
    u0 = 0.2 * torch.randn(n, 128, generator=g)
    uT = 0.1 * torch.randn(n, 128, generator=g)
    tgt = torch.randn(n, 128, generator=g)
    states = fx.calibration_states(n, seed=77)
    return u0, uT, tgt, states


def _worker(rank, ws, port, n_total, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    import safediffcon_b200 as s
    from safediffcon_b200 import runner
    from safediffcon_b200.conformal import kth_select, quantile_rank, scores_and_weights
    gd = _model()
    cfg = Cfg()
    u0, uT, tgt, states = _inputs(n_total)
    lo, hi = runner.shard_range(n_total, rank, ws)
    # ---- guided chains of the local shard (seed given, offset derived from the shard) + rollout + ONE all-gather ----
    pred = runner.sample_controls(gd, u0[lo:hi], uT[lo:hi], cfg, 0.0, n_total=n_total, seed=2024)
    metrics, _ = runner.evaluate_controls(pred, tgt[lo:hi], cfg.u_bound, n_total=n_total)
    traj, J, pts, tms, flg = s.control_and_score(pred, tgt[lo:hi].cuda(), cfg.u_bound, want_traj=False)
    # ---- calibration: scores of the local shard, all-gather, weights normalised over the full vector, k-th select ----
    q, sc, wn = runner.calibrate_quantile(gd, states[lo:hi], cfg, 0.02, 0.9, n_total=n_total, seed=99)
    out = dict(rank=rank, lo=lo, hi=hi, pred=pred.cpu(), metrics=metrics, J=J.cpu(), pts=pts.cpu(), tms=tms.cpu(), flg=flg.cpu(),
               q=q.cpu(), sc=sc.cpu(), wn=wn.cpu())
    if rank == 0:
        # the same job on ONE rank's worth of code: whole batch, offset 0, same seeds (no collective involved)
        full = gd.sample(batch_size=n_total, clip_denoised=True, u_init=u0.cuda(), u_final=uT.cuda(), guidance_u0=True,
                         nablaJ=s.safety_guidance(cfg, 0.0), enable_grad=False, seed=2024, sample_offset=0) * 10.0
        st = states.cuda()
        cal = gd.sample(batch_size=n_total, clip_denoised=True, guidance_u0=False, u_init=st[:, 0, 0, :], u_final=st[:, 0, 10, :],
                        w_groundtruth=st[:, 1], nablaJ=None, enable_grad=False, seed=99, sample_offset=0)
        sc1, w1 = scores_and_weights(cal, st, cfg, 0.02)
        wn1 = s.normalize_weights(w1.clone())
        out.update(full=full.cpu(), sc1=(wn1 * sc1).cpu(), wn1=wn1.cpu())   # weighted scores, what get_conformal_scores returns
    torch.save(out, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _check(tmp_path, ws, n_total):
    from safediffcon_b200 import runner
    from oracle import conformal_ref as cr
    mp.spawn(_worker, args=(ws, _free_port(), n_total, str(tmp_path)), nprocs=ws, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt", weights_only=False) for r in range(ws)]
    # every rank holds identical global results, bit for bit
    for r in res[1:]:
        assert r["metrics"] == res[0]["metrics"]
        assert torch.equal(r["q"], res[0]["q"]) and torch.equal(r["sc"], res[0]["sc"]) and torch.equal(r["wn"], res[0]["wn"])
    # gathered vectors = per-rank vectors in global index order; metrics = the reference formulae on them
    J = torch.cat([r["J"] for r in res]); pts = torch.cat([r["pts"] for r in res])
    tms = torch.cat([r["tms"] for r in res]); flg = torch.cat([r["flg"] for r in res])
    host = runner.metrics_from_vectors(J, pts.long(), tms.long(), flg.long(), 11, 128)   # same formulae on the host (CPU reduction order)
    for k, v in res[0]["metrics"].items():
        if k in ("control_mse_mean (J)", "control_mse_std"):
            assert abs(v - host[k]) <= 1e-6 * abs(host[k]), (k, v, host[k])
        else:
            assert v == host[k], (k, v, host[k])
    # quantile = the rank-th order statistic of the gathered scores (host sort), selected on the device
    sc = res[0]["sc"]
    assert sc.shape[0] == n_total
    assert res[0]["q"].item() == float(cr.quantile(sc, 0.9))
    # sharding does not change what is computed: per-rank chains == the single-rank chain over the whole batch
    full, pred = res[0]["full"], torch.cat([r["pred"] for r in res])
    assert torch.allclose(pred, full, rtol=1e-4, atol=1e-4), (pred - full).abs().max()
    assert torch.allclose(sc, res[0]["sc1"], rtol=1e-3, atol=1e-4), (sc - res[0]["sc1"]).abs().max()
    assert torch.allclose(res[0]["wn"], res[0]["wn1"], rtol=1e-3, atol=1e-5)
    print(f"world {ws}: n={n_total} bitwise-equal samples: {bool(torch.equal(pred, full))}, max |d| {(pred - full).abs().max().item():.2e}, "
          f"Q {res[0]['q'].item():.6f}, J {res[0]['metrics']['control_mse_mean (J)']:.6f}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("n_total", [12, 13])
def test_two_ranks_nccl(tmp_path, n_total):
    _check(tmp_path, 2, n_total)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs 8 GPUs (gpurun --gpus 8)")
def test_eight_ranks_nccl(tmp_path):
    _check(tmp_path, 8, 44)
