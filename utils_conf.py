"""Configuration files and command-line overrides for train.py / inference.py.

Behaviour follows the reference's utils_conf.py:4-42 so that its config files and `--set` strings keep
working: JSON unless the suffix says YAML, `--set a.b.c=value` with the value read as bool, int, float or
string, intermediate tables created on demand (a non-table in the way is replaced)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Iterable


def _read_yaml(fh) -> dict:
    try:
        import yaml
    except ImportError as exc:
        raise RuntimeError("PyYAML not installed; use a .json config or `pip install pyyaml`") from exc
    return yaml.safe_load(fh)


_READERS = {".yml": _read_yaml, ".yaml": _read_yaml}


def load_config(path: str) -> dict:
    """Parse a JSON (default) or YAML (.yml / .yaml) configuration file."""
    cfg_path = Path(path)
    if not cfg_path.exists():
        raise FileNotFoundError(f"Config not found: {cfg_path}")
    reader = _READERS.get(cfg_path.suffix.lower(), json.load)
    with cfg_path.open("r") as fh:
        return reader(fh)


def _coerce(text: str) -> Any:
    """'true' / 'false' -> bool; a number -> int, or float when it contains a dot; anything else stays text."""
    folded = text.lower()
    if folded == "true":
        return True
    if folded == "false":
        return False
    convert = float if "." in text else int
    try:
        return convert(text)
    except ValueError:
        return text


def apply_overrides(cfg: dict, pairs: Iterable[str]) -> None:
    """Apply `section.key=value` assignments in place, e.g. `train.batch_size=4 unet.base_ch=64`."""
    for assignment in pairs:
        dotted, sep, raw = assignment.partition("=")
        if not sep:
            raise ValueError(f"Invalid override (no '='): {assignment}")
        *tables, leaf = dotted.split(".")
        cursor = cfg
        for name in tables:
            if not isinstance(cursor.get(name), dict):
                cursor[name] = {}
            cursor = cursor[name]
        cursor[leaf] = _coerce(raw)
