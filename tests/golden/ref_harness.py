"""Import the UNMODIFIED reference (iitmdinesh/image2text) in this container.

Test infrastructure only: used by ``tests/golden/make_golden.py`` to produce the
committed golden fixtures, and by ``bench.py --impl reference`` / the ``cpu_baseline``
leg when a copy of the reference is reachable.  Nothing in ``image2text_b200/``
imports this file.

The reference needs two import shims (``peft`` and ``smart_open`` are not installed)
and offline patches (no hub weights): see SURVEY.md section 8(c) / Appendix A.
"""
import enum
import os
import sys
import types

REFERENCE_CANDIDATES = (
    os.environ.get("I2T_REFERENCE_ROOT", ""),
    os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "baseline", "_ref"),
    "/root/reference",
)


def find_reference_root():
    for cand in REFERENCE_CANDIDATES:
        if cand and os.path.isfile(os.path.join(cand, "models", "vision_encoder_decoder.py")):
            return cand
    return None


def install_shims():
    if "peft" not in sys.modules:
        peft = types.ModuleType("peft")
        tuners = types.ModuleType("peft.tuners")

        class TaskType(enum.Enum):
            FEATURE_EXTRACTION = "FEATURE_EXTRACTION"
            CAUSAL_LM = "CAUSAL_LM"

        class LoraConfig:
            def __init__(self, **kw):
                self.__dict__.update(kw)

        class LoraModel:  # only reached when lora_spec is not None
            def __init__(self, *a, **k):
                raise RuntimeError("peft is not installed in this image")

        peft.TaskType, peft.LoraConfig, peft.LoraModel = TaskType, LoraConfig, LoraModel
        peft.prepare_model_for_kbit_training = lambda m, **k: m
        tuners.LoraModel = LoraModel
        peft.tuners = tuners
        import importlib.machinery as _m
        peft.__spec__ = _m.ModuleSpec("peft", None)
        tuners.__spec__ = _m.ModuleSpec("peft.tuners", None)
        sys.modules["peft"], sys.modules["peft.tuners"] = peft, tuners
    if "smart_open" not in sys.modules:
        so = types.ModuleType("smart_open")
        so.open = open
        sys.modules["smart_open"] = so


_REF = {}


def load_reference():
    """Returns a namespace with the reference's modules, or raises if it is absent."""
    if _REF:
        return types.SimpleNamespace(**_REF)
    root = find_reference_root()
    if root is None:
        raise FileNotFoundError("reference checkout not found (looked in %r)" % (REFERENCE_CANDIDATES,))
    install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    import models.encoder as E  # noqa: E402  (reference module)
    import models.decoder as D  # noqa: E402
    _orig_vit = E.vit_b_16
    if not getattr(E.vit_b_16, "_i2t_offline", False):
        def _vit_offline(weights=None):
            return _orig_vit(weights=None)
        _vit_offline._i2t_offline = True
        E.vit_b_16 = _vit_offline
    from transformers import GPT2Config, GPT2LMHeadModel

    def _from_pretrained(model_str, config=None, **kw):
        return GPT2LMHeadModel(config if config is not None else GPT2Config())
    D.AutoModelForCausalLM.from_pretrained = staticmethod(_from_pretrained)
    import configs.models as CM  # noqa: E402
    import configs.trainer as CT  # noqa: E402
    import models.vision_encoder_decoder as VED  # noqa: E402
    import models.optimizer as OPT  # noqa: E402
    import models.utils as MU  # noqa: E402
    import training.wrapper as TW  # noqa: E402
    _REF.update(root=root, E=E, D=D, CM=CM, CT=CT, VED=VED, OPT=OPT, MU=MU, TW=TW)
    return types.SimpleNamespace(**_REF)


def fake_tokenizer(mask_token_id=None, vocab_size=50257, eos=50256, bos=50256):
    return types.SimpleNamespace(eos_token_id=eos, bos_token_id=bos,
                                 mask_token_id=mask_token_id, vocab_size=vocab_size)


def build_reference_model(model_cfg_dict, state_dict=None):
    """model_cfg_dict: the ``model:`` section of a training YAML (python dict)."""
    import copy
    ref = load_reference()
    d = copy.deepcopy(model_cfg_dict)
    dec = d["decoder_config"]
    if "pretrained_model" in dec:
        dec["pretrained_model"] = None  # random-init TransformerDecoder (decoder.py:44-46)
    if "lora_spec" in dec:
        dec["lora_spec"] = None
    if "lora_spec" in d["vision_encoder_config"]:
        d["vision_encoder_config"]["lora_spec"] = None
    cfg = ref.CM.VisionEncoderDecoderConfig.model_validate(d)
    model = ref.VED.VisionEncoderDecoder(cfg)
    if state_dict is not None:
        missing = model.load_state_dict(state_dict, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
    return model
