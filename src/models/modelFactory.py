"""Generator factory with the reference's call shape (reference src/models/modelFactory.py:6-8).
Only the AERO generator is provided by this repo; discriminators are outside the hot path."""
from aero_b200.model import Aero


def get_model(args):
    exp = args.experiment if hasattr(args, "experiment") else args["experiment"]
    model = exp.model if hasattr(exp, "model") else exp["model"]
    if model != "aero":
        raise NotImplementedError(f"aero_b200 provides the 'aero' generator only, got {model!r}")
    kw = exp.aero if hasattr(exp, "aero") else exp["aero"]
    return {"generator": Aero(**kw)}
