"""``add_enhance_arguments`` -- expose the keyword arguments of ``model.enhance`` on an argparse
parser (reference ``inference_utils/signature_to_parser.py:26-66``): the argument names and
casters come from the method's type hints (``Optional[T]`` -> ``T``), the defaults from
``model.diff_kwargs``."""
import typing


def add_enhance_arguments(model, parser):
    enhance = getattr(model, "enhance", None)
    if not callable(enhance):
        raise ValueError("Model does not have an `enhance` method.")
    hints = typing.get_type_hints(enhance)
    hints.pop("return", None)
    defaults = getattr(model, "diff_kwargs", {})
    group = parser.add_argument_group("enhance", "Arguments of enhance function")
    for key, hint in hints.items():
        inner = typing.get_args(hint)
        caster = inner[0] if inner else hint
        group.add_argument(f"--{key}", default=defaults.get(key, None), type=caster)
    return parser
