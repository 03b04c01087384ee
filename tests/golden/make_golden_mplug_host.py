"""Host-side golden values for the mPLUG driver helpers, from the UNMODIFIED reference packages mPLUG/scheduler and
mPLUG/optim on the miniature mPLUG-shaped network of tests/mplug_skeleton.py:

    python tests/golden/make_golden_mplug_host.py        # writes tests/golden/mplug_host.json
"""
import json
import os
import sys
import types


HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
sys.path.insert(0, os.path.join(os.environ.get("CRVQA_REFERENCE_ROOT", "/root/reference"), "mPLUG"))
import mplug_skeleton as sk  # noqa: E402

SCHED = dict(sched="cosine", lr=3e-5, epochs=8, min_lr=1e-6, decay_rate=1, warmup_lr=1e-5, warmup_epochs=4,
             cooldown_epochs=0)
OPT = dict(opt="adamW", lr1=3e-5, lr2=5e-6, lr=3e-5, weight_decay=0.02)


def groups(optimizer, model):
    names = {id(p): n for n, p in model.named_parameters()}
    vis = {id(p): "visual_encoder." + n for n, p in model.visual_encoder.named_parameters()}
    return [{"lr": g["lr"], "weight_decay": g["weight_decay"],
             "params": [names.get(id(p), vis.get(id(p))) for p in g["params"]]} for g in optimizer.param_groups]


def main():
    from optim import create_optimizer, create_two_optimizer      # the reference's packages
    from scheduler import create_scheduler
    out = {"sched": SCHED, "opt": OPT}
    model = sk.build()
    for n, p in model.named_parameters():                          # a masker-like trainable set
        p.requires_grad = ("predictions" in n) or n.endswith("intermediate.dense.weight")
    a = types.SimpleNamespace(**OPT)
    two = create_two_optimizer(a, model)
    out["two_optimizer_groups"] = groups(two, model)
    one = create_optimizer(a, model)
    out["optimizer_groups"] = groups(one, model)
    sch, epochs = create_scheduler(types.SimpleNamespace(**SCHED), two)
    out["num_epochs"] = epochs
    out["lr_after_init"] = [g["lr"] for g in two.param_groups]
    seq = []
    for t in list(range(0, 14)) + [3, 0, 20]:
        sch.step(t)
        seq.append([t, [g["lr"] for g in two.param_groups]])
    out["lr_by_step"] = seq
    sch2, _ = create_scheduler(types.SimpleNamespace(**dict(SCHED, warmup_epochs=0, lr_cycle_limit=2, lr_cycle_mul=2.0,
                                                            decay_rate=0.5, epochs=3)), one)
    seq = []
    for t in range(0, 12):
        sch2.step(t)
        seq.append([t, [g["lr"] for g in one.param_groups]])
    out["lr_by_step_cycles"] = seq
    out["cycle_length"] = sch2.get_cycle_length()
    with open(os.path.join(HERE, "mplug_host.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote mplug_host.json", epochs, out["lr_after_init"], seq[:3])


if __name__ == "__main__":
    main()
