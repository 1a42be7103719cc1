"""Drop-in glue for the reference's driver scripts.

``install()`` makes ``import tensorflow`` resolve to ``compat.tf`` and aliases the reference's
package paths (``models.ctc.ctc``, ``models.attention.*``, ``models.encoders.*``,
``utils.io.labels.sparsetensor``, ``utils.training.multi_gpu``) to this package, so a script
written against the reference imports the B200 implementation without edits.
"""
import importlib
import sys

_ALIASES = {
    "models": "models",
    "models.model_base": "models.model_base",
    "models.ctc": "models.ctc",
    "models.ctc.ctc": "models.ctc.ctc",
    "models.ctc.multitask_ctc": "models.ctc.multitask_ctc",
    "models.encoders.core.multitask_blstm": "models.encoders.core.multitask_blstm",
    "models.encoders": "models.encoders",
    "models.encoders.load_encoder": "models.encoders.load_encoder",
    "models.encoders.core": "models.encoders.core",
    "models.encoders.core.blstm": "models.encoders.core.blstm",
    "models.encoders.core.vgg_blstm": "models.encoders.core.vgg_blstm",
    "models.attention": "models.attention",
    "models.attention.attention_seq2seq": "models.attention.attention_seq2seq",
    "models.attention.joint_ctc_attention": "models.attention.joint_ctc_attention",
    "models.attention.bridge": "models.attention.bridge",
    "models.attention.decoders": "models.attention.decoders",
    "models.attention.decoders.attention_layer": "models.attention.decoders.attention_layer",
    "models.attention.decoders.attention_decoder": "models.attention.decoders.attention_decoder",
    "utils": "utils",
    "utils.io": "utils.io",
    "utils.io.labels": "utils.io.labels",
    "utils.io.labels.sparsetensor": "utils.io.labels.sparsetensor",
    "utils.evaluation": "utils.evaluation",
    "utils.evaluation.edit_distance": "utils.evaluation.edit_distance",
    "utils.training": "utils.training",
    "utils.training.multi_gpu": "utils.training.multi_gpu",
}


def install(alias_reference_packages=True, reference_root=None):
    """``reference_root``: checkout of the reference repository.  The aliased packages (``models``, ``utils``, ...)
    then also search the reference's own directories, so that modules this package does not replace --
    ``utils.training.learning_rate_controller``, ``utils.directory``, ``utils.parameter``,
    ``examples.*.metrics`` ... -- keep importing from the reference next to the B200 models."""
    from . import tf
    sys.modules["tensorflow"] = tf
    if alias_reference_packages:
        pkg = __name__.rsplit(".", 1)[0]
        for ref, mine in _ALIASES.items():
            m = importlib.import_module(pkg + "." + mine)
            sys.modules.setdefault(ref, m)
            if reference_root and hasattr(m, "__path__"):
                import os
                extra = os.path.join(reference_root, *ref.split("."))
                if os.path.isdir(extra) and extra not in list(m.__path__):
                    m.__path__.append(extra)
                    _EXTENDED.append((m, extra))
        if reference_root and reference_root not in sys.path:
            sys.path.append(reference_root)
            _EXTENDED.append((None, reference_root))
    return tf


_EXTENDED = []


def uninstall():
    sys.modules.pop("tensorflow", None)
    while _EXTENDED:
        m, extra = _EXTENDED.pop()
        if m is None:
            if extra in sys.path:
                sys.path.remove(extra)
        elif extra in list(m.__path__):
            m.__path__.remove(extra)
    for name in [n for n in sys.modules if n == "examples" or n.startswith("examples.")]:
        sys.modules.pop(name, None)
    pkg = __name__.rsplit(".", 1)[0]
    for ref in _ALIASES:
        m = sys.modules.get(ref)
        if m is not None and getattr(m, "__name__", "").startswith(pkg):
            sys.modules.pop(ref, None)
