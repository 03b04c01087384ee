"""Optimiser construction of the mPLUG driver (reference mPLUG/optim/optim_factory.py): ``create_optimizer`` (the
shipped ``opt: adamW``: torch AdamW over the TRAINABLE parameters, 1-D tensors and biases without weight decay, :31-45,
:60-89) and ``create_two_optimizer`` (:141-171: separate learning rates for the text side and the visual encoder).
The other optimiser families of that factory (and its vendored implementations) are not built."""
from torch import optim


def add_weight_decay(model, weight_decay=1e-5, skip_list=()):
    decay, no_decay = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue                                # frozen weights: under the masker, everything but scores + LM head
        if len(param.shape) == 1 or name.endswith(".bias") or name in skip_list:
            no_decay.append(param)
        else:
            decay.append(param)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


_FAMILIES = {"adamw": optim.AdamW, "adam": optim.Adam}


def create_optimizer(args, model, filter_bias_and_bn=True):
    opt_lower = args.opt.lower()
    weight_decay = args.weight_decay
    if weight_decay and filter_bias_and_bn:
        skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else {}
        parameters = add_weight_decay(model, weight_decay, skip)
        weight_decay = 0.0
    else:
        parameters = model.parameters()
    opt_args = dict(lr=args.lr, weight_decay=weight_decay)
    if getattr(args, "opt_eps", None) is not None:
        opt_args["eps"] = args.opt_eps
    if getattr(args, "opt_betas", None) is not None:
        opt_args["betas"] = args.opt_betas
    if getattr(args, "opt_args", None) is not None:
        opt_args.update(args.opt_args)
    family = opt_lower.split("_")[-1]
    if family not in _FAMILIES or "_" in opt_lower:
        raise NotImplementedError(f"opt={args.opt!r}: only adamW / adam are built")
    return _FAMILIES[family](parameters, **opt_args)


def create_two_optimizer(args, model, filter_bias_and_bn=True):
    """Four groups: (decay, no-decay) x (everything outside ``visual_encoder`` at lr1, ``model.visual_encoder`` at lr2).
    As in the reference, the visual groups filter on names RELATIVE to ``model.visual_encoder`` with the test
    ``"visual_encoder" not in name`` -- which those relative names always pass -- and frozen tensors are included
    (torch skips parameters without a gradient)."""
    no_decay = ["bias", "LayerNorm.weight"]

    def pick(named, want_no_decay):
        return [p for n, p in named if any(nd in n for nd in no_decay) == want_no_decay and "visual_encoder" not in n]

    groups = [
        {"params": pick(model.named_parameters(), False), "weight_decay": args.weight_decay, "lr": args.lr1},
        {"params": pick(model.named_parameters(), True), "weight_decay": 0.0, "lr": args.lr1},
        {"params": pick(model.visual_encoder.named_parameters(), False), "weight_decay": args.weight_decay,
         "lr": args.lr2},
        {"params": pick(model.visual_encoder.named_parameters(), True), "weight_decay": 0.0, "lr": args.lr2},
    ]
    return optim.AdamW(groups)
