#!/usr/bin/env python
"""TEST INFRASTRUCTURE (not product code).  Composes the reference's Hydra configuration tree the way
``make_model`` does (experiments/benchmark_utils.py:159-171: ``compose(overrides=["+target=..", "+solver=..",
"model@generative_ctrl=..", "loss.method=.."])``) with a minimal re-statement of Hydra's defaults-list rules (hydra and
omegaconf are not installed here) and stores the resolved, flattened trees in tests/golden/conf_defaults.json:

    python -m oracle.make_conf_golden          # reads /root/reference/conf (this container only)

tests/test_make_model_rules.py holds ``benchmark_utils.default_config`` (the product's restatement of these YAMLs) to it.
Rules implemented (the subset conf/ uses): defaults entries ``name`` (same group), ``/group: name``, ``/group@pkg: name``,
``group@pkg: name`` (relative), ``_self_``; the ``# @package _global_`` header; later entries win; ``${a.b}``
interpolations resolved at the end."""
import json
import os
import re
import sys

import yaml

CONF = "/root/reference/conf"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "conf_defaults.json")


def merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            merge(dst[k], v)
        else:
            dst[k] = v if not isinstance(v, dict) else merge({}, v)
    return dst


def place(tree, package, content):
    node = tree
    for key in [k for k in package.split(".") if k]:
        node = node.setdefault(key, {})
    merge(node, content)


def load(group, name, package, tree, overrides):
    """Merges conf/<group>/<name>.yaml (and its defaults) into ``tree`` at ``package``."""
    path = os.path.join(CONF, group, name + ".yaml")
    text = open(path).read()
    if re.search(r"#\s*@package\s+_global_", text):
        package = ""
    body = yaml.safe_load(text) or {}
    defaults = body.pop("defaults", ["_self_"])
    if "_self_" not in defaults:
        defaults = ["_self_"] + defaults  # Hydra 1.1+: _self_ first when absent
    for entry in defaults:
        if entry == "_self_":
            place(tree, package, body)
        elif isinstance(entry, str):
            load(group, entry, package, tree, overrides)
        else:
            (key, val), = entry.items()
            val = overrides.get(key.lstrip("/"), val)
            if val is None:  # "- /sde:" (conf/solver/dds.yaml): the group is left empty
                continue
            grp, _, pkg = key.partition("@")
            absolute = grp.startswith("/")
            grp = grp.lstrip("/")
            sub_group = grp if absolute else os.path.join(group, grp)
            sub_pkg = pkg if pkg else grp.split("/")[-1]
            if not absolute or package:  # relative packages nest under the including file's package
                sub_pkg = (package + "." if package else "") + sub_pkg
            load(sub_group, val, sub_pkg, tree, overrides)


def resolve(tree):
    def get(path):
        node = tree
        for k in path.split("."):
            node = node[k]
        return node

    def walk(node):
        for k, v in list(node.items()):
            if isinstance(v, dict):
                walk(v)
            elif isinstance(v, str):
                m = re.fullmatch(r"\$\{([\w.]+)\}", v)
                if m:
                    node[k] = get(m.group(1))
    for _ in range(3):
        walk(tree)


def flatten(node, prefix="", out=None):
    out = {} if out is None else out
    for k, v in node.items():
        key = prefix + k
        if isinstance(v, dict):
            flatten(v, key + ".", out)
        else:
            if k == "_target_":
                v = v.split(".")[-1]
            if isinstance(v, str):
                try:
                    v = float(v)  # YAML 1.1 reads 1e8 / 1e4 as strings
                except ValueError:
                    pass
            out[key] = v
    return out


def compose(solver, model, target, loss_method, sde=None):
    tree = {}
    overrides = {"model@generative_ctrl": model}
    if sde:
        overrides["sde"] = sde
    load("target", target, "target", tree, overrides)
    load("solver", solver, "", tree, overrides)
    tree["loss"]["method"] = loss_method
    resolve(tree)
    return flatten(tree)


def main():
    golden = {}
    for solver in ("vp_rds", "pbm_rds", "pis", "dds", "cmcd", "dis"):
        for model in ("basic", "score", "langevin_init", "lerp"):
            for method in ("lv", "kl"):
                golden[f"{solver}|{model}|{method}|many_modes|"] = compose(solver, model, "many_modes", method)
    for solver in ("vp_rds", "dis"):
        for sde in ("vp_20", "vp_cos"):
            golden[f"{solver}|score|lv|many_modes|{sde}"] = compose(solver, "score", "many_modes", "lv", sde)
    for target in ("two_modes", "bracket_two_modes", "phi_four", "sonar", "ionosphere", "cancer", "credit"):
        golden[f"vp_rds|score|lv|{target}|"] = compose("vp_rds", "score", target, "lv")
    json.dump(golden, open(OUT, "w"), indent=0, sort_keys=True)
    print(f"wrote {OUT}: {len(golden)} composed trees, {sum(len(v) for v in golden.values())} leaves")


if __name__ == "__main__":
    sys.exit(main())
