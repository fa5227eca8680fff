"""Generates the REFERENCE-PINNED golden fixtures tests/golden/ref_*.npz.

Unlike make_golden.py (outputs of the oracle), every array written here is an output of the reference's OWN source
files, imported unmodified from /root/reference/{cld_jax,blur_jax} under tests/refshim (a numpy / torch-CPU stand-in
for jax 0.2.8, flax 0.3.1, jammy, ml_collections; fp64 = the reference's `x64=True` mode).  Executed reference code:
  cld_jax/deis.py, sde_lib.py (CLD, LambdaSDE, LSDE), sampling.py (all sampler factories, MLCLD), utils.py,
  models/{utils,ncsnpp,layerspp,layers,up_or_down_sampling}.py,
  blur_jax/sde_lib.py, blur.py, fft.py, sampling.py, multistep.py, models/*.
/root/reference only exists in the authoring container; the fixtures are committed so that the CPU and GPU tests
(tests/test_ref_golden.py) can check the oracle, the host tables and the CUDA path against them anywhere.

  python tests/golden/make_ref_golden.py            # both trees (a few minutes of python loops)
  python tests/golden/make_ref_golden.py cld|blur   # one tree
"""
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

inv = lambda x: (x + 1.) / 2.   # noqa: E731  datasets.py:34-40 with data.centered


def nest(flat):
  out = {}
  for k, v in flat.items():
    node = out
    parts = k.split("/")
    for q in parts[:-1]:
      node = node.setdefault(q, {})
    node[parts[-1]] = np.asarray(v, np.float64)
  return out


def small_net(mutils, refshim, jnp, kind, cld=True):
  """The reference NCSNpp on tests/helpers.small_cfg(kind); parameters from the product's name-keyed generator, over the
  (name, shape, initializer) list the REFERENCE module creates (not the oracle's walk)."""
  from helpers import small_cfg
  from gddim_b200 import params
  cfg = small_cfg(kind)
  model = mutils.get_model("ncsnpp")(config=cfg)
  cin = 6 if cld else 3
  specs = refshim.linen_collect(model, jnp.zeros((1, 32, 32, cin)), jnp.ones(1), train=False)
  specs = {k: (s, kd, (1e-10 if (kd == "vs" and sc == 0) else sc)) for k, (s, kd, sc) in specs.items()}
  p = params.generate(specs, seed=1234, nondegenerate=True)
  return cfg, model, specs, p


def spec_arrays(specs):
  names = np.array(list(specs.keys()))
  shapes = np.array(["x".join(str(d) for d in s) for s, _, _ in specs.values()])
  kinds = np.array([k for _, k, _ in specs.values()])
  scales = np.array([sc for _, _, sc in specs.values()], np.float64)
  return dict(names=names, shapes=shapes, kinds=kinds, scales=scales)


def gen_cld():
  import refshim
  jax = refshim.install("cld_jax")
  import jax.numpy as jnp
  import deis
  import sampling
  import sde_lib
  from models import up_or_down_sampling as uds
  from models import utils as mutils
  from models import ncsnpp  # noqa: F401  (registers the model, as run_lib.py does)
  from helpers import prior_u
  t00 = time.time()
  want = lambda name: not SECTIONS or name in SECTIONS   # noqa: E731
  out = {}

  # ---- 1. CLD tables: sde_lib.py:45-319, deis.py:19-95 -------------------------------------------------------------
  t_at = np.array([1e-3, 0.0123, 0.1, 0.37, 0.5, 0.9, 1.0])
  for tag, kw in (("rk", dict(is_R_rk=True, R_dt=1e-3)), ("euler", dict(is_R_rk=False, R_dt=1e-3)),
                  ("b1", dict(is_R_rk=True, R_dt=1e-3, beta_0=0.5, beta_1=3.0, m_inv=2.0, vv_gamma=0.02))) if want("tables") else ():
    sde = sde_lib.CLD(used_cache=False, x64=True, **kw)
    out[f"{tag}_R"] = np.stack([sde.s_R(t) for t in t_at])
    out[f"{tag}_psi"] = np.stack([sde.s_psi(s, t) for s, t in zip(t_at[1:], t_at[:-1])])
    out[f"{tag}_F"] = np.stack([sde.s_F(t) for t in t_at])
    out[f"{tag}_G"] = np.stack([sde.s_G(t) for t in t_at])
    out[f"{tag}_integrand"] = np.stack([sde.s_eps_integrand(t) for t in t_at])
    rev = sampling.get_rev_ts(sde, 2, 9)
    out[f"{tag}_rev_ts"] = rev
    for order in ((0, 1, 2, 3) if tag == "rk" else (2,)):
      out[f"{tag}_deis_o{order}"] = sde.get_deis_coef(order, rev)
      print(f"[cld] {tag} get_deis_coef order {order}  t={time.time() - t00:.0f}s", flush=True)
    m, e = sde.prepare_order0_coef(rev)
    out[f"{tag}_order0_mean"], out[f"{tag}_order0_eps"] = m, e
    m, e = sde.prepare_naive_coef(rev)
    out[f"{tag}_naive_mean"], out[f"{tag}_naive_eps"] = m, e
    if tag == "rk":
      eps = np.random.default_rng(5).standard_normal((2, 4, 4, 3, 2))
      out["rk_eps2score_in"] = eps
      out["rk_eps2score"] = sde.eps2score(jnp.asarray(eps), jnp.asarray([0.3, 0.7]))
  out["t_at"] = t_at
  if want("tables"):
    np.savez_compressed(os.path.join(HERE, "ref_cld_tables.npz"), **{k: np.asarray(v) for k, v in out.items()})
  print(f"[cld] tables done t={time.time() - t00:.0f}s", flush=True)
  if want("variants"):
    gen_cld_variants(sde_lib, sampling, jnp, t00)
  if want("ops"):
    gen_cld_ops(deis, uds, jnp, t00)
  if want("net") or want("samplers"):
    gen_cld_net_samplers(refshim, jax, jnp, mutils, sde_lib, sampling, prior_u, t00, want("samplers"))


def gen_cld_variants(sde_lib, sampling, jnp, t00):

  # ---- 2. LambdaSDE / LSDE / MLCLD tables: sde_lib.py:334-519, sampling.py:272-316 ----------------------------------
  out = {}
  sde = sde_lib.CLD(used_cache=False, x64=True, is_R_rk=False, R_dt=1e-4)        # the "cld_mixed" test table
  rev5 = sampling.get_rev_ts(sde, 2, 4)
  out["rev_ts"] = rev5
  lam = sde_lib.LambdaSDE(sde, 0.5, True, used_cache=False)
  print(f"[cld] LambdaSDE hat_psi table t={time.time() - t00:.0f}s", flush=True)
  out["lambda05_order0_coef"] = lam.get_order0_coef(rev5, used_cache=False)
  out["lambda05_deis_o0"] = lam.get_deis_coef(0, rev5, used_cache=False)
  out["lambda05_deis_o1"] = lam.get_deis_coef(1, rev5, used_cache=False)
  out["lambda05_hat_psi"] = np.stack([lam.s_hat_psi(s, t) for s, t in zip(rev5[:-1], rev5[1:])])
  print(f"[cld] LambdaSDE tables t={time.time() - t00:.0f}s", flush=True)
  ls = sde_lib.LSDE(sde, used_cache=False)
  rev6 = sampling.get_rev_ts(sde, 2, 6)
  out["rev_ts6"] = rev6
  out["lsde_deis_o2"] = ls.get_deis_coef(2, rev6, used_cache=False)
  out["lsde_L"] = np.stack([ls.s_L(t) for t in rev6])
  e = np.random.default_rng(6).standard_normal((2, 3, 3, 2))
  out["lsde_epsR2epsL_in"], out["lsde_epsR2epsL"] = e, ls.epsR2epsL(0.4, jnp.asarray(e))
  ml = sampling.MLCLD(sde)
  rev5d = sampling.get_rev_ts(sde, 2, 5)
  out["rev_ts5"] = rev5d
  out["mlcld_deis_o1"] = ml.get_deis_coef(1, rev5d)
  out["mlcld_psi2"] = np.stack([ml.s_psi2_fn(t) for t in rev5d])
  print(f"[cld] MLCLD tables t={time.time() - t00:.0f}s", flush=True)
  np.savez_compressed(os.path.join(HERE, "ref_cld_variants.npz"), **{k: np.asarray(v) for k, v in out.items()})



def gen_cld_ops(deis, uds, jnp, t00):
  # ---- 3. update operator + resamplers: deis.py:141-151, up_or_down_sampling.py:76-86,168-411 -----------------------
  out = {}
  rng = np.random.default_rng(7)
  for order in range(4):
    x = rng.standard_normal((3, 8, 8, 3, 2))
    ne = rng.standard_normal((3, 8, 8, 3, 2))
    hist = rng.standard_normal((order + 1, 3, 8, 8, 3, 2))
    coef = rng.standard_normal((order + 3, 2, 2))
    xn, hn = deis.multistep_ab_step(jnp.asarray(x), jnp.asarray(coef), jnp.asarray(ne), jnp.asarray(hist))
    out.update({f"ab{order}_x": x, f"ab{order}_new_eps": ne, f"ab{order}_hist": hist, f"ab{order}_coef": coef,
                f"ab{order}_x_next": xn, f"ab{order}_hist_next": hn})
  img = rng.standard_normal((2, 8, 8, 64)).astype(np.float32).astype(np.float64)
  w = rng.standard_normal((3, 3, 64, 4))
  out["fir_in"], out["fir_w"] = img, w
  out["fir_up"] = uds.upsample_2d(jnp.asarray(img), (1, 3, 3, 1), factor=2)
  out["fir_down"] = uds.downsample_2d(jnp.asarray(img), (1, 3, 3, 1), factor=2)
  out["fir_conv_down"] = uds.conv_downsample_2d(jnp.asarray(img), jnp.asarray(w), k=(1, 3, 3, 1))
  out["naive_up"] = uds.naive_upsample_2d(jnp.asarray(img))
  out["naive_down"] = uds.naive_downsample_2d(jnp.asarray(img))
  np.savez_compressed(os.path.join(HERE, "ref_ops.npz"), **{k: np.asarray(v) for k, v in out.items()})
  print(f"[cld] operators done t={time.time() - t00:.0f}s", flush=True)



def gen_cld_net_samplers(refshim, jax, jnp, mutils, sde_lib, sampling, prior_u, t00, do_samplers):
  # ---- 4. NCSN++ / DDPM++ forward: models/ncsnpp.py:41-243 and the layer files ---------------------------------------
  out = {}
  nets = {}
  for kind in ("cld_deep", "cld_ddpmpp"):
    cfg, model, specs, p = small_net(mutils, refshim, jnp, kind)
    nets[kind] = (cfg, model, p)
    out.update({f"{kind}_spec_{k}": v for k, v in spec_arrays(specs).items()})
    x = np.random.default_rng(8).standard_normal((2, 32, 32, 6)).astype(np.float32).astype(np.float64)
    out[f"{kind}_x"] = x.astype(np.float32)
    for j, t in enumerate((0.37, 0.004)):
      y = model.apply({"params": nest(p)}, jnp.asarray(x), jnp.ones(2) * 999.0 * t, train=False, mutable=False)
      out[f"{kind}_t{j}"], out[f"{kind}_y{j}"] = t, np.asarray(y)
  np.savez_compressed(os.path.join(HERE, "ref_net.npz"), **out)
  print(f"[cld] network forward done t={time.time() - t00:.0f}s", flush=True)

  if not do_samplers:
    return
  # ---- 5. samplers end to end: sampling.py:41-669, models/utils.py:128-182 -------------------------------------------
  out = {}

  def state_of(p):
    return mutils.State(step=0, optimizer=None, lr=0.0, model_state={}, ema_rate=0.0, params_ema=nest(p), rng=None)

  def cld_of(cfg, mixed=None):
    m = cfg.model
    return sde_lib.CLD(m_inv=m.m_inv, beta_0=m.beta_0, beta_1=m.beta_1, vv_gamma=m.vv_gamma,
                       mixed_score=m.mixed_score if mixed is None else mixed, is_R_rk=m.is_R_rk, used_cache=False,
                       R_dt=m.R_dt, x64=True)

  def run(tag, fn, state, u, batch):
    refshim.clear_draws()
    r = fn(jax.random.PRNGKey(0), state, batch, jnp.asarray(u.astype(np.float64)))
    out[f"{tag}_u"] = u
    out[f"{tag}_x"], out[f"{tag}_v"], out[f"{tag}_nfe"] = np.asarray(r[0]), np.asarray(r[1]), int(r[2])
    if refshim.draws():
      out[f"{tag}_z"] = np.stack(refshim.draws()).astype(np.float32)
    print(f"[cld] sampler {tag} nfe={int(r[2])} t={time.time() - t00:.0f}s", flush=True)

  from helpers import small_cfg
  cfg_m = small_cfg("cld_mixed")                               # Euler R table, R_dt = 1e-4
  _, model, p = nets["cld_deep"]                               # same architecture and parameters as "cld_mixed"
  st = state_of(p)
  sde = cld_of(cfg_m, mixed=False)
  shape = (32, 32, 3)
  run("deis_o2", sampling.get_deis_sampler(sde, model, shape, 6, inv, 2, ts_order=2, denoising=True), st, prior_u(2, seed=0), 2)
  run("deis_o3", sampling.get_deis_sampler(sde, model, shape, 8, inv, 3, ts_order=2, denoising=True), st, prior_u(2, seed=1), 2)
  run("deis_o0_nodenoise", sampling.get_deis_sampler(sde, model, shape, 5, inv, 0, ts_order=2, denoising=False), st, prior_u(3, seed=2), 3)
  run("order0", sampling.get_order0_sampler(sde, model, shape, 6, inv, is_em=False, denoising=True), st, prior_u(2, seed=3), 2)
  run("order0_em", sampling.get_order0_sampler(sde, model, shape, 6, inv, is_em=True, denoising=True), st, prior_u(2, seed=68), 2)
  # through get_sampling_fn + psampler (leading device axis), hybdeis time grid
  cfg_m.sampling.method, cfg_m.sampling.nfe, cfg_m.sampling.deis_order = "hybdeis", 9, 1
  fn = sampling.get_sampling_fn(cfg_m, sde, model, None, inv)
  u = prior_u(2, seed=31)
  r = fn(jax.random.PRNGKey(0)[None], sys.modules["flax.jax_utils"].replicate(st), 2, jnp.asarray(u[None].astype(np.float64)))
  out["hybdeis_u"], out["hybdeis_x"], out["hybdeis_v"], out["hybdeis_nfe"] = u, np.asarray(r[0]), np.asarray(r[1]), int(r[2])
  print(f"[cld] sampler hybdeis (psampler) t={time.time() - t00:.0f}s", flush=True)
  # (denoising=True raises AttributeError in the reference: LambdaSDE has no sampling_eps / s_F, sampling.py:383)
  run("sdeis", sampling.get_sdeis_sampler(sde, model, shape, 5, inv, 1, lambda_coef=0.5, use_order0=True, ts_order=2, denoising=False),
      st, prior_u(2, seed=41), 2)
  run("ldeis", sampling.get_L_deis_sampler(sde, model, shape, 6, inv, 2, ts_order=2, denoising=False), st, prior_u(2, seed=61), 2)
  run("em", sampling.get_em_sampler(sde, model, shape, 6, inv, lambda_coef=0.7, ts_order=2, denoising=True), st, prior_u(2, seed=62), 2)
  run("sscs", sampling.get_sscs_sampler(sde, model, shape, 5, inv, ts_order=2, denoising=False), st, prior_u(2, seed=64), 2)
  run("mldeis", sampling.get_mldeis_sampler(sde, model, shape, 6, inv, 1, ts_order=2, denoising=True), st, prior_u(2, seed=67), 2)
  run("mixed_deis_o2", sampling.get_deis_sampler(cld_of(cfg_m, mixed=True), model, shape, 6, inv, 2, ts_order=2, denoising=True),
      st, prior_u(2, seed=4), 2)
  # DDPM++ (positional embedding, naive resampling, no pyramid)
  cfg_d, model_d, p_d = nets["cld_ddpmpp"]
  cfg_d.model.R_dt = 1e-4
  run("ddpmpp_deis_o1", sampling.get_deis_sampler(cld_of(cfg_d), model_d, shape, 6, inv, 1, ts_order=2, denoising=True),
      state_of(p_d), prior_u(2, seed=5), 2)
  np.savez_compressed(os.path.join(HERE, "ref_cld_samplers.npz"), **{k: np.asarray(v) for k, v in out.items()})
  print(f"[cld] all done t={time.time() - t00:.0f}s", flush=True)


def gen_blur():
  import refshim
  jax = refshim.install("blur_jax")
  import jax.numpy as jnp
  import blur
  import multistep
  import sampling
  import sde_lib
  from models import utils as mutils
  from models import ncsnpp  # noqa: F401  (registers the model, as run_lib.py does)
  from helpers import prior_u
  t00 = time.time()
  out = {}
  # ---- blur SDE schedule: blur_jax/sde_lib.py:18-163, sampling.py:42-51 ----------------------------------------------
  for smax in (1.0, 10.0):
    s = sde_lib.SDE(sigma_blur_max=smax, sampling_eps=1e-5)
    tag = f"s{int(smax)}"
    rev = sampling.get_rev_ts(s, 2, 50)
    out[f"{tag}_rev_ts"] = rev
    out[f"{tag}_sampling_T"] = float(s.sampling_T)
    ts = jnp.asarray(np.asarray(rev)[[0, 1, 7, 25, 49, 50]])
    out[f"{tag}_ts"] = ts
    out[f"{tag}_alpha"] = s.t2alpha_fn(ts)
    out[f"{tag}_freq_scaling"] = s.get_frequency_scaling(ts)
    out[f"{tag}_y_mean_coef"] = s.y_mean_coef(ts)
    out[f"{tag}_y_std_coef"] = s.y_std_coef(ts)
    out[f"{tag}_rho2t"] = np.array([float(s.rho2t(r)) for r in (0.5, 7.0, 80.0)])
  # ---- DCT / IDCT: blur.py:11-107 (Makhoul FFT trick over fft.py -> lax.fft) ----------------------------------------
  x = np.random.default_rng(11).standard_normal((3, 32, 32, 3))
  out["dct_in"] = x
  out["dct"] = blur.batch_img_dct(jnp.asarray(x))
  out["idct"] = blur.batch_img_idct(jnp.asarray(x))
  # ---- scalar-coefficient AB update: multistep.py:94-98 --------------------------------------------------------------
  rng = np.random.default_rng(12)
  for order in range(3):
    xx, ne = rng.standard_normal((2, 6, 6, 3)), rng.standard_normal((2, 6, 6, 3))
    hist, coef = rng.standard_normal((order + 1, 2, 6, 6, 3)), rng.standard_normal(order + 3)
    xn, hn = multistep.ab_step(jnp.asarray(xx), jnp.asarray(coef), jnp.asarray(ne), jnp.asarray(hist))
    out.update({f"sab{order}_x": xx, f"sab{order}_new_eps": ne, f"sab{order}_hist": hist, f"sab{order}_coef": coef,
                f"sab{order}_x_next": xn, f"sab{order}_hist_next": hn})
  np.savez_compressed(os.path.join(HERE, "ref_blur_tables.npz"), **{k: np.asarray(v) for k, v in out.items()})
  print(f"[blur] tables/DCT done t={time.time() - t00:.0f}s", flush=True)
  # ---- network (C_in = 3) + order-0 sampler: blur_jax/sampling.py:53-90, models/utils.py:104-160 ---------------------
  out = {}
  cfg, model, specs, p = small_net(mutils, refshim, jnp, "blur_deep", cld=False)
  out.update({f"blur_deep_spec_{k}": v for k, v in spec_arrays(specs).items()})
  xin = np.random.default_rng(13).standard_normal((2, 32, 32, 3)).astype(np.float32)
  out["net_x"], out["net_t"] = xin, 0.61
  out["net_y"] = model.apply({"params": nest(p)}, jnp.asarray(xin.astype(np.float64)), jnp.ones(2) * 999.0 * 0.61, train=False, mutable=False)
  st = mutils.State(step=0, optimizer=None, lr=0.0, model_state={}, ema_rate=0.0, params_ema=nest(p), rng=None)
  s = sde_lib.from_config(cfg)
  y = prior_u(2, seed=2, cld=False)
  fn = sampling.get_order0_sampler(s, model, (32, 32, 3), 2, 6, inv, is_p=False)
  xs, nfe = fn(jax.random.PRNGKey(0), st, 2, jnp.asarray(y.astype(np.float64)))
  out["order0_y"], out["order0_x"], out["order0_nfe"] = y, np.asarray(xs), int(nfe)
  cfg.sampling.nfe = 5
  pfn = sampling.get_sampling_fn(cfg, s, model, None, inv)
  xs, nfe = pfn(jax.random.PRNGKey(0)[None], sys.modules["flax.jax_utils"].replicate(st), 2, jnp.asarray(y[None].astype(np.float64)))
  out["p_order0_x"], out["p_order0_nfe"] = np.asarray(xs), int(nfe)
  np.savez_compressed(os.path.join(HERE, "ref_blur_sampler.npz"), **{k: np.asarray(v) for k, v in out.items()})
  print(f"[blur] all done t={time.time() - t00:.0f}s", flush=True)


SECTIONS = set(sys.argv[2:])      # cld only: tables variants ops net samplers (default: all)

if __name__ == "__main__":
  if len(sys.argv) > 1:
    {"cld": gen_cld, "blur": gen_blur}[sys.argv[1]]()
  else:                       # both trees use the same top-level module names -> one process each
    for tree in ("cld", "blur"):
      subprocess.check_call([sys.executable, os.path.abspath(__file__), tree])
