"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py

Every rank holds a contiguous row shard (ensemble.row_partition) of the same seeded global dataset; the
scalars returned by the C ABI (all-reduced over NCCL by the library) must equal the oracle on the FULL
data, and each rank's per-row outputs must equal the oracle's rows of its shard."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from spark_ensemble_b200 import _native as N
    from spark_ensemble_b200.context import Context
    from spark_ensemble_b200.ensemble import row_partition
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    uid = torch.zeros(N.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(Context.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    assert ctx.comm_info() == (world, rank)
    p2p = ctx.comm_p2p_active()
    if os.environ.get("SE_P2P_ALLREDUCE", "1") != "0" and os.environ.get("SE_REQUIRE_P2P") == "1":
        assert p2p, "fused peer-memory all-reduce expected to be active"
    orc = O.Oracle()
    rng = np.random.default_rng(11)
    n, nv, K = 200_003, 50_001, 5
    RT = 1e-5
    for name in ("squared", "bernoulli", "logloss"):
        dim = K if name == "logloss" else 1
        lid = O.LOSS_IDS[name]
        y = (rng.integers(0, K, n) if name == "logloss" else rng.standard_normal(n) if name == "squared"
             else (rng.random(n) < 0.4)).astype(np.float32)
        vy = (rng.integers(0, K, nv) if name == "logloss" else rng.standard_normal(nv) if name == "squared"
              else (rng.random(nv) < 0.4)).astype(np.float32)
        w = (rng.random(n) + 0.5).astype(np.float32)
        F = (0.5 * rng.standard_normal((dim, n))).astype(np.float32)
        h = rng.standard_normal((dim, n)).astype(np.float32)
        vF = (0.5 * rng.standard_normal((dim, nv))).astype(np.float32)
        s0, s1 = row_partition(n, world, rank)
        v0, v1 = row_partition(nv, world, rank)
        ctx.gbm_configure(s1 - s0, v1 - v0, dim, name, 0.0, True)
        ctx.upload(N.SLOT_Y, y[s0:s1]); ctx.upload(N.SLOT_W, w[s0:s1])
        ctx.upload(N.SLOT_F, np.ascontiguousarray(F[:, s0:s1])); ctx.upload(N.SLOT_H, np.ascontiguousarray(h[:, s0:s1]))
        ctx.upload(N.SLOT_VY, vy[v0:v1]); ctx.upload(N.SLOT_VF, np.ascontiguousarray(vF[:, v0:v1]))
        alpha = rng.random(dim) + 0.5
        lg, gg = ctx.gbm_linesearch_eval(alpha)
        lo, go = orc.linesearch_eval(lid, 0.0, y, w, F, h, alpha)
        assert abs(lg - lo) <= RT * abs(lo), (name, lg, lo)
        assert np.all(np.abs(gg - go) <= RT * np.maximum(np.abs(go), np.abs(go).max())), (name, gg, go)
        assert abs(ctx.gbm_mean_loss(True) - orc.mean_loss(lid, 0.0, dim, vy, vF)) <= RT * abs(orc.mean_loss(lid, 0.0, dim, vy, vF))
        S = ctx.gbm_pseudo_residuals(newton=True)
        ro, wo, So = orc.pseudo_residuals(lid, 0.0, dim, y, w, F, True)
        assert np.all(np.abs(S - So) <= RT * So), (name, S, So)
        wg = ctx.download(N.SLOT_WOUT).reshape(dim, -1)
        assert np.all(np.abs(wg - wo[:, s0:s1]) <= RT * np.abs(wo[:, s0:s1]) + 1e-12), name
        step = rng.random(dim) * 0.5
        ls, _ = ctx.gbm_update(step, residual=True, loss=True)
        Fo = F.astype(np.float64).copy()
        orc.update(Fo, h, step)
        lo = orc.mean_loss(lid, 0.0, dim, y, Fo) * n
        assert abs(ls - lo) <= RT * abs(lo), (name, ls, lo)
        Fg = ctx.download(N.SLOT_F).reshape(dim, -1)
        assert np.max(np.abs(Fg - Fo[:, s0:s1])) <= RT * max(1.0, np.abs(Fo).max())
        if name == "squared":
            st = ctx.gbm_linesearch_stats()
            d = y.astype(np.float64) - Fo[0]
            ref = [np.sum(d * d), np.sum(h[0] * d), np.sum(h[0].astype(np.float64) ** 2), np.sum(w.astype(np.float64))]
            assert np.all(np.abs(st - ref) <= RT * np.abs(ref) + 1e-6), (st, ref)
            ctx.gbm_round_squared_async(0.5)
            a, lsum = ctx.gbm_round_result()
            star = float(np.clip(ref[1] / ref[2], 0, 100))
            assert abs(a - star) <= RT * max(abs(star), 1e-3), (a, star)
    # SAMME.R across shards
    Kc = 7
    yb = rng.integers(0, Kc, n).astype(np.float32)
    Z = rng.standard_normal((Kc, n)); Z[yb.astype(int), np.arange(n)] += 1.5
    P = (np.exp(Z) / np.exp(Z).sum(0)).astype(np.float32)
    wb = (rng.random(n) + 0.1).astype(np.float32)
    s0, s1 = row_partition(n, world, rank)
    ctx.boost_configure(s1 - s0, Kc, True)
    ctx.upload(N.SLOT_Y, yb[s0:s1]); ctx.upload(N.SLOT_BW, wb[s0:s1]); ctx.upload(N.SLOT_PROBA, np.ascontiguousarray(P[:, s0:s1]))
    sw = ctx.slot_sum(N.SLOT_BW)
    assert abs(sw - orc.sum(wb)) <= 1e-9 * sw
    e, s = ctx.boost_real_update(sw)
    out, eo, so = orc.samme_r_update(Kc, yb, wb, sw, P)
    assert abs(e - eo) <= RT * eo and abs(s - so) <= RT * so, (e, eo, s, so)
    assert np.max(np.abs(ctx.download(N.SLOT_BW) - out[s0:s1]) / out[s0:s1]) <= RT
    # exact quantile (radix select, histograms all-reduced) and AdaBoost.R2's max-allreduce across shards
    from spark_ensemble_b200.ensemble import exact_quantile
    v = rng.standard_normal(n).astype(np.float32)
    ctx.alloc(N.SLOT_Y, s1 - s0)
    ctx.upload(N.SLOT_Y, v[s0:s1])
    for q in (0.1, 0.5, 0.9):
        assert ctx.quantile(N.SLOT_Y, q) == exact_quantile(v, q)
    pr = (v + 0.3 * rng.standard_normal(n)).astype(np.float32)
    ctx.boostreg_configure(s1 - s0)
    ctx.upload(N.SLOT_Y, v[s0:s1]); ctx.upload(N.SLOT_PRED, pr[s0:s1]); ctx.upload(N.SLOT_BW, wb[s0:s1])
    mx = ctx.boostreg_max_error()
    assert abs(mx - orc.r2_max_error(v, pr)) <= 1e-6 * mx
    e2 = ctx.boostreg_error(sw, "exponential", mx)
    assert abs(e2 - orc.r2_estimator_error("exponential", v, pr, wb, sw, mx)) <= RT * e2
    # ---- LogLoss beyond 64 classes: the general kernels, sums all-reduced by NCCL across the shards
    Kb, nb = 70, 30_011
    yb = rng.integers(0, Kb, nb).astype(np.float32)
    Fb = (0.5 * rng.standard_normal((Kb, nb))).astype(np.float32)
    hb = rng.standard_normal((Kb, nb)).astype(np.float32)
    s0, s1 = row_partition(nb, world, rank)
    ctx.gbm_configure(s1 - s0, 0, Kb, "logloss", 0.0, False)
    ctx.upload(N.SLOT_Y, yb[s0:s1]); ctx.upload(N.SLOT_F, np.ascontiguousarray(Fb[:, s0:s1])); ctx.upload(N.SLOT_H, np.ascontiguousarray(hb[:, s0:s1]))
    ab = rng.random(Kb) + 0.5
    lg, gg = ctx.gbm_linesearch_eval(ab)
    lo, go = orc.linesearch_eval(O.LOGLOSS, 0.0, yb, None, Fb, hb, ab)
    assert abs(lg - lo) <= RT * abs(lo), (lg, lo)
    assert np.all(np.abs(gg - go) <= RT * np.maximum(np.abs(go), np.abs(go).max()))
    # ---- an EMPTY local validation shard still joins the validation collectives
    nv_tiny = world - 1  # one validation row on every rank but the last
    yv = rng.standard_normal(max(nv_tiny, 1)).astype(np.float32)
    Fv = rng.standard_normal(max(nv_tiny, 1)).astype(np.float32)
    mine = 1 if rank < world - 1 else 0
    ctx.gbm_configure(1000, mine, 1, "squared", 0.0, False)
    ctx.fill(N.SLOT_Y, 1.0); ctx.fill(N.SLOT_F, 0.0); ctx.fill(N.SLOT_H, 1.0)
    if mine:
        ctx.upload(N.SLOT_VY, yv[rank:rank + 1]); ctx.upload(N.SLOT_VF, Fv[rank:rank + 1]); ctx.fill(N.SLOT_VH, 0.0)
    ml = ctx.gbm_mean_loss(True)
    ref = float(np.mean(0.5 * (yv[:nv_tiny].astype(np.float64) - Fv[:nv_tiny]) ** 2))
    assert abs(ml - ref) <= RT * ref, (ml, ref)
    assert abs(ctx.gbm_update_validation([0.0]) - ref) <= RT * ref
    # ---- cooperative kernels across shards: the cross-GPU sums happen INSIDE the running kernel (peer mailboxes)
    import time
    if p2p:
        n2 = 1_000_003
        for name in ("squared", "bernoulli", "absolute"):
            lid = O.LOSS_IDS[name]
            y = (rng.standard_normal(n2) if name != "bernoulli" else (rng.random(n2) < 0.4)).astype(np.float32)
            F = (0.5 * rng.standard_normal((1, n2))).astype(np.float32)
            r0, _, _ = orc.pseudo_residuals(lid, 0.0, 1, y, None, F, False)
            h = (0.6 * r0 + 0.2 * rng.standard_normal((1, n2))).astype(np.float32)
            s0, s1 = row_partition(n2, world, rank)
            ctx.gbm_configure(s1 - s0, 0, 1, name, 0.0, False)
            ctx.upload(N.SLOT_Y, y[s0:s1]); ctx.upload(N.SLOT_F, np.ascontiguousarray(F[:, s0:s1]))
            ctx.upload(N.SLOT_H, np.ascontiguousarray(h[:, s0:s1]))
            f = lambda x: orc.linesearch_eval(lid, 0.0, y, None, F, h, [x])[0]
            ao, neo, st = orc.brent(f)
            assert st == 0
            if name == "squared":
                ctx.set_option("fused_round", 1)
                a, ls, ne = ctx.gbm_round(0.5, True, 1e-6, 100, residual=True)
                assert ctx.get_option("last_round_fused") == 1
                ctx.set_option("fused_round", -1)
                assert abs(a - ao) <= 1e-5 * abs(ao) + 4e-6, (a, ao)
                Fo = F.astype(np.float64).copy()
                orc.update(Fo, h, [0.5 * a])
                lo = orc.mean_loss(lid, 0.0, 1, y, Fo) * n2
                assert abs(ls - lo) <= RT * abs(lo), (ls, lo)
                assert np.max(np.abs(ctx.download(N.SLOT_F) - Fo[0, s0:s1])) <= RT * max(1.0, np.abs(Fo).max())
                # every rank must hold the SAME alpha bit for bit (rank-ordered sums in the mailboxes)
                t = torch.tensor([a], dtype=torch.float64, device="cuda")
                tl = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(tl, t)
                assert all(float(x) == a for x in tl), [float(x) for x in tl]
            else:
                ctx.set_option("ls_mode", 1)
                dev = ctx.gbm_linesearch_brent()
                ctx.set_option("ls_mode", 2)
                host = ctx.gbm_linesearch_brent()
                ctx.set_option("ls_mode", 1)
                assert dev == host, (name, dev, host)
                assert abs(dev[1] - f(ao)) <= RT * abs(f(ao)), (name, dev, f(ao))
                assert f(dev[0]) <= f(ao) * (1 + 1e-5)
        # ---- skewed ranks: a late rank inside the (configurable) bound is simply waited for ...
        ctx.alloc(N.SLOT_BW, 1000)
        ctx.fill(N.SLOT_BW, 1.0)
        ctx.set_option("peer_timeout_ms", 20000.0)
        dist.barrier()
        if rank == world - 1:
            time.sleep(4.0)  # longer than round 1's hard-coded 3 s bound
        assert ctx.slot_sum(N.SLOT_BW) == 1000.0 * world
        # ... and one beyond it fails the SAME reduction on every rank (poisoned mailbox rows), then recovers
        ctx.set_option("peer_timeout_ms", 300.0)
        dist.barrier()
        if rank == world - 1:
            time.sleep(2.0)
        failed = False
        try:
            ctx.slot_sum(N.SLOT_BW)
        except N.NativeError as e:
            failed = "peer-memory all-reduce failed" in str(e)
        assert failed, "a reduction whose peer never showed up in time must fail on every rank"
        dist.barrier()
        ctx.comm_clear_error()
        ctx.set_option("peer_timeout_ms", 20000.0)
        dist.barrier()
        assert ctx.slot_sum(N.SLOT_BW) == 1000.0 * world
    dist.barrier()
    if rank == 0:
        print(f"MGPU_PARITY_OK world={world} p2p_allreduce={p2p}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
