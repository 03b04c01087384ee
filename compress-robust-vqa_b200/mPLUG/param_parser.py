"""``k=v,k=v`` option strings (reference mPLUG/param_parser.py:7-27; the argparse Action classes of that file are
not used by the mask-training path)."""


def str2bool(v):
    low = v.lower()
    if low in ("yes", "true", "t", "y", "1"):
        return True
    if low in ("no", "false", "f", "n", "0"):
        return False
    return v


def dict_parser(values):
    """Numbers become floats, yes/no words booleans, anything else stays a string."""
    parsed = {}
    for item in values.split(","):
        key, val = item.split("=")
        try:
            parsed[key] = float(val)
        except ValueError:
            parsed[key] = str2bool(val)
    return parsed
