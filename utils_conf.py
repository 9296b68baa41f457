"""Config loading with `--set a.b=c` overrides (same semantics as the reference's utils_conf.py:4-42):
JSON by default, YAML only for .yml/.yaml, values parsed as bool / int / float / string, missing
intermediate dicts created."""
import json
import pathlib


def load_config(path: str) -> dict:
    p = pathlib.Path(path)
    if not p.exists():
        raise FileNotFoundError(f"Config not found: {p}")
    if p.suffix.lower() in (".yml", ".yaml"):
        try:
            import yaml
        except ImportError as e:
            raise RuntimeError("PyYAML not installed; use a .json config or `pip install pyyaml`") from e
        with p.open("r") as f:
            return yaml.safe_load(f)
    with p.open("r") as f:
        return json.load(f)


def _parse_value(s: str):
    low = s.lower()
    if low in ("true", "false"):
        return low == "true"
    try:
        return float(s) if "." in s else int(s)
    except ValueError:
        return s


def apply_overrides(cfg: dict, pairs) -> None:
    for pair in pairs:
        if "=" not in pair:
            raise ValueError(f"Invalid override (no '='): {pair}")
        key, val = pair.split("=", 1)
        node = cfg
        parts = key.split(".")
        for k in parts[:-1]:
            if k not in node or not isinstance(node[k], dict):
                node[k] = {}
            node = node[k]
        node[parts[-1]] = _parse_value(val)
