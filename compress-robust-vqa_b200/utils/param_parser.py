"""``k=v,k=v`` option strings (reference utils/param_parser.py:6-26): used for masking_scheduler_conf."""


def str2bool(v):
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    return v


def dict_parser(values):
    parsed = {}
    for kv in values.split(","):
        k, v = kv.split("=")
        try:
            parsed[k] = float(v)
        except ValueError:
            parsed[k] = str2bool(v)
    return parsed
