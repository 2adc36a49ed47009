"""Probe-head construction for --cls_features ep / ep_all, mirroring the reference's probe_heads.py.

    model.head = Sequential(EfficientProbing, BatchNorm1d(width, affine=False, eps=1e-6), Linear)

Only the EP row of the reference's POOLINGS table (probe_heads.py:66-84) is in scope; the thirteen
other attentive poolings are out of scope for this repository and raise NotImplementedError.  Names
with no pooling (cls, gap, ...) get the plain BatchNorm + classifier probe exactly as
probe_heads.py:98-101 does."""
import torch.nn as nn

from .ep import EfficientProbing

_OTHER_POOLINGS = ("abmilp", "simpool", "esimpool", "clip", "siglip", "aim", "cbam", "coca", "cait", "dinovit",
                   "jepa", "dolg", "cae")

# name -> (pooling factory, classifier factory)      probe_heads.py:75-76
POOLINGS = {
    "ep": (lambda dim, a, m: EfficientProbing(dim=dim, num_queries=a.ep_queries, d_out=a.d_out),
           lambda dim, a: nn.Linear(dim // a.d_out, a.nb_classes, bias=True)),
}


def _batchnorm(width):
    return nn.BatchNorm1d(width, affine=False, eps=1e-6)          # probe_heads.py:109-110


def build_probe_head(model, args):
    """Replace model.head in place with the probe selected by args.cls_features (probe_heads.py:87-106)."""
    name = args.cls_features
    base = name[:-len("_all")] if name.endswith("_all") else name                # probe_heads.py:95
    dim = model.head.in_features
    if base in _OTHER_POOLINGS:
        raise NotImplementedError(f"--cls_features {name}: only the EP pooling is implemented here")
    if base not in POOLINGS:
        model.head = nn.Sequential(_batchnorm(dim), model.head)                  # plain linear probe
        return
    make_pooling, make_classifier = POOLINGS[base]
    pooling = make_pooling(dim, args, model)          # built first: fixes the RNG order (probe_heads.py:104)
    classifier = make_classifier(dim, args)
    model.head = nn.Sequential(pooling, _batchnorm(classifier.in_features), classifier)


def make_ep_head(dim, num_queries=32, nb_classes=1000, d_out=1, qkv_bias=False):
    """The EP probe head without a surrounding encoder (what training on cached tokens needs)."""
    pooling = EfficientProbing(dim=dim, num_queries=num_queries, d_out=d_out, qkv_bias=qkv_bias)
    classifier = nn.Linear(dim // d_out, nb_classes, bias=True)
    return nn.Sequential(pooling, _batchnorm(classifier.in_features), classifier)
