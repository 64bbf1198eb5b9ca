import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import restate as R
from oracle.golden import load_golden
from oracle.make_golden import toy_model_fn
from sdb200.ddim import DDIMSampler
g = load_golden("ddim.pt")
class Shim(R.ModelShim):
    def apply_model(self, x, tt, c):
        return toy_model_fn(x.cpu(), tt.cpu(), c.cpu()).cuda()
for S, cfg in ((10, 1.0), (50, 1.0), (10, 5.0)):
    t = g["traj.S%d.cfg%g" % (S, cfg)]
    orc = R.DDIMOracle(R.ModelShim(toy_model_fn, R.sd_alphas_cumprod())); orc.make_schedule(S)
    shim = Shim(None, R.sd_alphas_cumprod(), device="cuda"); shim.betas = shim.betas.cuda()
    smp = DDIMSampler(shim); smp.make_schedule(S, verbose=False)
    x = t["x_T"]; c = t["c"]; uc = t["uc"]
    bad = 0
    for i, step in enumerate(np.flip(orc.ddim_timesteps)):
        index = S - i - 1
        ts = torch.full((3,), int(step), dtype=torch.long)
        xo, p0o, e = orc.p_sample_ddim(x, c, ts, index, unconditional_guidance_scale=cfg, unconditional_conditioning=uc)
        xg, p0g = smp.p_sample_ddim(x.cuda(), c.cuda(), ts.cuda(), index=index, unconditional_guidance_scale=cfg,
                                    unconditional_conditioning=None if uc is None else uc.cuda())
        dx = (xg.cpu() != xo).sum().item(); dp = (p0g.cpu() != p0o).sum().item()
        if dx or dp:
            bad += 1
            if bad <= 2:
                # is the eps itself different (batched 2B call vs two B calls on the CPU toy model)?
                eu = toy_model_fn(x, ts, uc) if uc is not None else None
                ec = toy_model_fn(x, ts, c)
                e2 = ec if eu is None else eu + cfg * (ec - eu)
                print("  S%d cfg%g index %d: x_prev mism %d pred_x0 mism %d ; eps(batched) vs eps(separate) mism %d" %
                      (S, cfg, index, dx, dp, (e2 != e).sum().item()))
        x = xo
    zg, _ = DDIMSampler(shim).sample(S, 3, (4, 8, 8), conditioning=c.cuda(), verbose=False, x_T=t["x_T"].cuda(), eta=0.,
                                     unconditional_guidance_scale=cfg, unconditional_conditioning=None if uc is None else uc.cuda())
    zo, _ = R.DDIMOracle(R.ModelShim(toy_model_fn, R.sd_alphas_cumprod())).sample(S, 3, (4, 8, 8), conditioning=c, eta=0., x_T=t["x_T"],
                                     unconditional_guidance_scale=cfg, unconditional_conditioning=uc)
    print("S%d cfg%g: teacher-forced bad steps %d ; free-running mismatches vs oracle %d, vs golden %d (oracle vs golden %d)" %
          (S, cfg, bad, (zg.cpu() != zo).sum().item(), (zg.cpu() != t["z"]).sum().item(), (zo != t["z"]).sum().item()))
