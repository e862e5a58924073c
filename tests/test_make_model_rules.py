"""CPU: make_model / make_target_details keep the reference's selectors, defaults and forbidden combinations
(experiments/benchmark_utils.py:41-159; conf/solver/*.yaml - SURVEY.md Appendix A).  Nothing here touches a GPU."""
import pytest

from sde_sampler_lrds_b200 import benchmark_utils as BU
from sde_sampler_lrds_b200.losses import oc as L
from sde_sampler_lrds_b200.solver import oc as S

TRAIN = {"train_steps": 16, "train_batch_size": 8, "eval_batch_size": 8}


def mk(**kw):
    args = dict(solver_type="vp-ref", ref_type="default", loss_type="lv", integrator_type="ei",
                model_type="target_informed_zero_init", time_type="uniform", solver_details={"sigma": 1.0},
                target_details=BU.make_target_details("two_modes"), training_details=TRAIN)
    args.update(kw)
    return BU.make_model(**args)


@pytest.mark.parametrize("kw, msg", [
    (dict(solver_type="pis_orig", model_type="base_zero_init", integrator_type="em"), "Only target_informed_zero_init"),
    (dict(solver_type="pis_orig", integrator_type="em", time_type="snr"), "Only uniform time"),
    (dict(solver_type="dds_orig", integrator_type="ei"), "Can't use EI or DDPM-like discretization"),
    (dict(solver_type="pbm-ref", time_type="uniform"), "PBM schedule is unstable"),
    (dict(integrator_type="ddpm_like", time_type="uniform"), "DDPM with uniform times"),
    (dict(solver_type="pis_orig", integrator_type="em", ref_type="gmm"), "Only ref models"),
    (dict(solver_type="cmcd", integrator_type="em", model_type="base_zero_init", force_base_zero_init=True,
          ref_type="gmm"), "Can't use ref other than gaussian"),
    (dict(force_vp20=True, force_vp_cosine=True), "at the same time"),
    (dict(model_type="target_informed_lerp_tempering"), "lerp_tempering is not supported"),
])
def test_forbidden_combinations_raise_like_the_reference(kw, msg):
    with pytest.raises(ValueError, match=msg):
        mk(**kw)


def test_target_details_defaults_including_the_reference_typo():
    d = BU.make_target_details("two_modes")
    assert d == {"name": "two_modes", "dim": 5, "ill_conditioned": "medium", "a": 1.0}
    assert BU.make_target_details("many_modes", dim=50, n_modes=16)["n_modes"] == 16
    assert BU.make_target_details("phi_four") == {"name": "phi_four", "dim": 100, "b": 0.0}
    assert BU.make_target_details("sonar") == {"name": "sonar"}


def test_default_config_resolves_the_yaml_defaults():
    c = BU._resolve(BU.default_config("vp-ref", "target_informed_zero_init", "lv", BU.make_target_details("many_modes")))
    assert c["solver"] is S.RDS and c["loss"]["_target_"] is L.EMReferenceSDELoss and c["loss"]["max_rnd"] == 1e8
    assert c["sde"]["diff_coeff_sq_max"] == 10.0 and c["prior"]["scale"] == 1.0
    assert c["train_timesteps"]["end"] == 1.0 and c["train_timesteps"]["steps"] == 100
    c = BU._resolve(BU.default_config("dds_orig", "target_informed_zero_init", "lv", BU.make_target_details("phi_four")))
    assert c["solver"] is S.DDS and c["sde"] is None and c["train_timesteps"]["dt"] == 0.05 and c["train_timesteps"]["end"] == 6.4
    assert c["prior"]["scale"] == 1.0 and c["target"]["beta"] == 20.0 and c["target"]["dim"] == 100
    c = BU._resolve(BU.default_config("cmcd", "base_zero_init", "lv", BU.make_target_details("two_modes")))
    assert c["solver"] is S.CMCD and c["prior"]["scale"] == 5.0 and c["loss"]["max_rnd"] is None and c["sde"]["clip_score"] == 1e5
    c = BU._resolve(BU.default_config("pbm-ref", "base_zero_init", "lv", BU.make_target_details("two_modes")))
    assert c["train_timesteps"]["start"] == 1e-4 and c["train_timesteps"]["end"] == 5.0
    c = BU._resolve(BU.default_config("dis_orig", "target_informed_zero_init", "lv", BU.make_target_details("two_modes"),
                                      force_vp20=True))
    assert c["solver"] is S.Bridge and c["loss"]["_target_"] is L.TimeReversalLoss and c["loss"]["max_rnd"] == 1e8
    assert c["sde"]["diff_coeff_sq_max"] == 20.0 and c["prior"]["scale"] == 1.0 and c["train_timesteps"]["end"] == 1.0


# ---- default_config against the reference's composed YAML tree ------------------------------------------------------
# tests/golden/conf_defaults.json = the trees Hydra composes for make_model's overrides (experiments/benchmark_utils.py:
# 159-171), produced from /root/reference/conf by oracle/make_conf_golden.py.  Keys of the control plane that the product
# does not carry (optimiser, logging / evaluation cadence, Sinkhorn, ...) are listed, not silently skipped.
CONTROL_PLANE = ("optim.", "eval_freq", "eval_stddev_steps", "eval_interval", "eval_device", "eval_init",
                 "ckpt_interval", "log_interval", "max_loss", "max_grad", "scale_loss")


def _flatten(node, prefix="", out=None):
    out = {} if out is None else out
    for k, v in node.items():
        if k.startswith("_wants"):  # product-internal wiring flags
            continue
        key = prefix + k
        if isinstance(v, dict):
            _flatten(v, key + ".", out)
        else:
            if k in ("_target_", "solver") and not isinstance(v, str):
                v = v.__name__
            out[key if k != "solver" else "solver._target_"] = v
    return out


def _golden():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "conf_defaults.json")))


SOLVER_KEY = {v: k for k, v in BU.solver_types.items()}
MODEL_KEY = {"basic": "base_zero_init", "score": "target_informed_zero_init", "langevin_init": "target_informed_langevin_init",
             "lerp": "target_informed_lerp_tempering"}


@pytest.mark.parametrize("key", sorted(_golden()))
def test_default_config_equals_the_composed_reference_yaml(key):
    solver, model, method, target, sde = key.split("|")
    want = {k: v for k, v in _golden()[key].items() if not k.startswith(CONTROL_PLANE)}
    cfg = BU.default_config(SOLVER_KEY[solver], MODEL_KEY[model], method, {"name": target}, force_vp20=sde == "vp_20",
                            force_vp_cosine=sde == "vp_cos")
    BU._resolve(cfg)
    got = _flatten(cfg)
    if want.get("target.data_dir") is None:
        got.pop("target.data_dir", None), want.pop("target.data_dir", None)
    if cfg.get("sde") is None:
        got.pop("sde", None)
    got = {k: v for k, v in got.items() if k not in ("seed", "device")}  # conf/base.yaml, not part of the solver tree
    missing = sorted(set(want) - set(got))
    extra = sorted(set(got) - set(want))
    assert not missing and not extra, (missing, extra)
    diff = {k: (got[k], want[k]) for k in want if got[k] != want[k] and not (got[k] is None and want[k] is None)}
    assert not diff, diff
