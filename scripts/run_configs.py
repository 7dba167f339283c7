"""BASELINE.json configs 3, 4 and 5 on 1..8 B200 (one process per GPU; launch with torchrun for N > 1):

  --config 3  burgers_numeric_solve rollout + J / safety scoring of 100k synthetic trajectories, sharded
  --config 4  conformal calibration: unguided w_groundtruth chain on --units calibration states (default 50k), nonconformity
              scores, all-gather, on-device weight normalisation and rank-th order statistic (alpha 0.98)
  --config 5  guided sampling, 1024 control instances per GPU, Q from --q (config 4's output), then rollout + metrics

Each prints ONE JSON line on rank 0 (device-timed with CUDA events, max over ranks).  Inputs are synthetic
(safediffcon_b200.synthetic, per-rank seeds derived from the global sample index ranges), weights are the seed-42
random initialisation, the sampler is the reference default DDIM-200 (eta 1) unless --ddpm is given.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class Cfg:
    use_max_safety = True
    u_bound = 0.8
    guidance_weights = {"w_score": 500.0}
    nt = 11
    InfFT_Q = None


def setup():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return int(os.environ["RANK"]), ws
    torch.cuda.set_device(0)
    return 0, 1


def max_ranks(x, ws):
    if ws == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def barrier(ws):
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()


def build_model(steps, ddpm):
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
    return s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=1000 if ddpm else steps,
                               ddim_sampling_eta=1.0, temporal=True, use_conv2d=True, is_condition_u0=True, is_condition_uT=True,
                               condition_idx=10, train_on_padded_locations=False).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4, 5])
    ap.add_argument("--units", type=int, default=None, help="total units (config 3: 100000 trajectories, config 4: 50000 states)")
    ap.add_argument("--batch", type=int, default=1024, help="chain batch per GPU")
    ap.add_argument("--steps", type=int, default=200, help="DDIM sampling steps")
    ap.add_argument("--ddpm", action="store_true", help="full DDPM-1000 chain instead of DDIM")
    ap.add_argument("--q", type=float, default=0.0, help="conformal quantile fed to the guidance (config 5)")
    ap.add_argument("--alpha", type=float, default=0.98)
    args = ap.parse_args()
    rank, ws = setup()
    import safediffcon_b200 as s
    from safediffcon_b200 import runner
    from safediffcon_b200.synthetic import burgers_instances
    cfg = Cfg()
    runner.all_gather_concat(torch.zeros(4, 2, device="cuda"), 4 * ws)   # NCCL communicator set-up is not part of any timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    line = {"config": args.config, "n_gpus": ws, "data": "synthetic", "weights": "seed-42 random init"}

    if args.config == 3:
        n = args.units or 100000
        lo, hi = runner.shard_range(n, rank, ws)
        u0, f = burgers_instances(hi - lo, seed=300 + rank)
        tgt, _ = burgers_instances(hi - lo, seed=900 + rank)
        u0_d, f_d, tgt_d = (torch.from_numpy(a).cuda() for a in (u0, f, tgt))
        for _ in range(2):
            traj = s.burgers_numeric_solve_free(u0_d, f_d, 0.01, 1.0)
            s.burgers_score(traj, tgt_d, cfg.u_bound)
        barrier(ws)
        e0.record()
        traj = s.burgers_numeric_solve_free(u0_d, f_d, 0.01, 1.0)
        J, pts, tms, flg = s.burgers_score(traj, tgt_d, cfg.u_bound)
        packed = torch.stack([J, pts.float(), tms.float(), flg.float()], dim=1)
        g = runner.all_gather_concat(packed, n)
        e1.record()
        barrier(ws)
        ms = max_ranks(e0.elapsed_time(e1), ws)
        m = runner.metrics_from_vectors(g[:, 0], g[:, 1].long(), g[:, 2].long(), g[:, 3].long(), 11, 128)
        line.update(metric="burgers_rollouts_per_s", value=n / (ms / 1e3), unit="rollouts/s", n=n, ms=ms, mode="strict fp32",
                    fp32_tflops_algorithmic=n * 17.92e6 / (ms / 1e3) / 1e12, finite=bool(torch.isfinite(traj).all().item()),
                    J=m["control_mse_mean (J)"], R_p=m["point_exceed_ratio (R_p)"], R_s=m["sample_exceed_ratio (R_s)"])

    elif args.config == 4:
        n = args.units or 50000
        lo, hi = runner.shard_range(n, rank, ws)
        gd = build_model(args.steps, args.ddpm)
        t_host0 = time.perf_counter()
        u0, f = burgers_instances(hi - lo, seed=400 + rank)
        f_d = torch.from_numpy(f).cuda()
        traj = s.burgers_numeric_solve_free(torch.from_numpy(u0).cuda(), f_d, 0.01, 1.0)
        states = s.dataset_states(traj, f_d).cpu().pin_memory()   # (u, f, s) / 10, s = max u^2 broadcast; calibration states start on the host
        t_data = time.perf_counter() - t_host0
        barrier(ws)
        e0.record()
        scores, weights = [], []
        for b0 in range(0, hi - lo, args.batch):
            st = states[b0:b0 + args.batch].cuda(non_blocking=True)
            pred = gd.sample(batch_size=st.shape[0], clip_denoised=True, guidance_u0=False, u_init=st[:, 0, 0, :],
                             u_final=st[:, 0, cfg.nt - 1, :], w_groundtruth=st[:, 1, :, :], nablaJ=None, enable_grad=False,
                             seed=4040, sample_offset=lo + b0)
            sc, w = s.conformal.scores_and_weights(pred, st, cfg, args.q)
            scores.append(sc)
            weights.append(w)
        e_mid = torch.cuda.Event(enable_timing=True)
        e_mid.record()
        from safediffcon_b200 import _lib as L
        g = runner.all_gather_concat(torch.stack([torch.cat(scores), torch.cat(weights)], dim=1), n)
        sc, w = g[:, 0].contiguous(), g[:, 1].contiguous()
        wn = torch.empty_like(w)
        L.check(L.lib().sdc_normalize_weights(L.ptr(w), L.ptr(wn), L.ptr(sc), n, L.stream_ptr()))
        weighted = (sc * wn).contiguous()          # what ConformalCalculator.get_conformal_scores returns (conformal.py:84)
        rk = s.conformal.quantile_rank(n, args.alpha)
        q, idx = s.conformal.kth_select(weighted, rk)
        e1.record()
        barrier(ws)
        ms = max_ranks(e0.elapsed_time(e1), ws)
        ms_tail = max_ranks(e_mid.elapsed_time(e1), ws)
        # bit-exact check of the selection against a host sort of the same gathered vector
        ref = torch.sort(weighted.cpu()).values[rk].item()
        line.update(metric="calibration_samples_per_s", value=n / (ms / 1e3), unit="samples/s", n=n, ms=ms, sampler="ddpm1000" if args.ddpm
                    else f"ddim{args.steps}", rank_selected=rk, quantile=q.item(), quantile_matches_host_sort=bool(q.item() == ref),
                    gather_normalise_select_ms=ms_tail, host_data_prep_s=t_data, alpha=args.alpha)

    else:
        B = args.batch
        gd = build_model(args.steps, args.ddpm)
        u0, _ = burgers_instances(B, seed=1000 + rank)
        tgt, _ = burgers_instances(B, seed=5000 + rank)
        u0_h, uT_h, tgt_h = (torch.from_numpy(a).pin_memory() for a in (u0 / 10.0, tgt / 10.0, tgt))
        runner.sample_controls(gd, u0_h[:8], uT_h[:8], cfg, args.q, sample_offset=0, seed=1)   # warm-up: packing, FiLM table
        barrier(ws)
        e0.record()
        pred = runner.sample_controls(gd, u0_h, uT_h, cfg, args.q, sample_offset=rank * B, seed=2024)
        m, _ = runner.evaluate_controls(pred, tgt_h, cfg.u_bound, n_total=ws * B)
        e1.record()
        barrier(ws)
        ms = max_ranks(e0.elapsed_time(e1), ws)
        line.update(metric="guided_chain_samples_per_s", value=ws * B / (ms / 1e3), unit="samples/s", n=ws * B, ms=ms,
                    sampler="ddpm1000" if args.ddpm else f"ddim{args.steps}", Q=args.q, J=m["control_mse_mean (J)"],
                    R_p=m["point_exceed_ratio (R_p)"], R_t=m["time_exceed_ratio (R_t)"], R_s=m["sample_exceed_ratio (R_s)"])

    if rank == 0:
        print(json.dumps(line), flush=True)
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
