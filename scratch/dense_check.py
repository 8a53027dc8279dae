import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from rspl_slam_b200 import capi, synth
from rspl_slam_b200.problem import LocalBatch
from oracle import orc
ctx = capi.Context(0)
probs = [synth.make_local_problem(synth.config_seed(5, 1), n_kf=40, n_points=5000, n_lines=500, loops=1),
         synth.make_local_problem(synth.config_seed(5, 2), n_kf=48, n_points=4000, n_lines=300)]
batch = LocalBatch.from_problems(probs)
res = ctx.local_batch(batch)
for w, p in enumerate(probs):
    fine = p.copy(); cfg6 = orc.make_config(None, numeric_delta=1e-6); orc.local_ba(fine, cfg6)
    ref = p.copy(); orc.local_ba(ref)
    a, b = batch.line_begin[w], batch.line_begin[w + 1]
    df = np.abs(res.line_wd[:, a:b].T - fine.line_L).max(axis=1)
    dr = np.abs(ref.line_L - fine.line_L).max(axis=1)
    print(w, 'lines vs fine: median %.2e q90 %.2e q99 %.2e max %.2e' % (np.median(df), np.quantile(df, .9), np.quantile(df, .99), df.max()))
    print(w, 'oracle faithful vs fine: median %.2e q90 %.2e max %.2e' % (np.median(dr), np.quantile(dr, .9), dr.max()))
    a, b = batch.point_begin[w], batch.point_begin[w + 1]
    dp = np.abs(res.point_xyz[:, a:b].T - fine.point_p).max(axis=1)
    print(w, 'points vs fine: median %.2e q99 %.2e max %.2e' % (np.median(dp), np.quantile(dp, .99), dp.max()))
    a, b = batch.pose_begin[w], batch.pose_begin[w + 1]
    print(w, 'pose dp max %.2e' % np.abs(res.pose_twc[:3, a:b].T - ref.pose_p).max(), 'vs fine %.2e' % np.abs(res.pose_twc[:3, a:b].T - fine.pose_p).max())
# timing of a single large window
for kf, npt, nln in ((40, 5000, 500), (100, 20000, 2000), (200, 50000, 5000)):
    p = synth.make_local_problem(synth.config_seed(5, 10), n_kf=kf, n_points=npt, n_lines=nln, loops=1)
    bb = LocalBatch.from_problems([p])
    ctx.local_batch_upload(bb)
    ctx.local_batch_solve(); ctx.sync()
    t0 = time.perf_counter(); ctx.local_batch_solve(); ctx.sync(); dt = time.perf_counter() - t0
    r = ctx.local_batch_download(ctx.alloc_local_result(bb))
    print(kf, 'KF', npt, 'pts: solve %.1f ms' % (dt * 1e3), 'iters', r.stats['iters'][0][:2], 'trials', r.stats['trials'][0][:2], 'chi2 %.1f' % r.stats['final_chi2'][0])
