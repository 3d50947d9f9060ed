"""YAML config loader with the reference's interface (yaml_parser.py:3-33):
``YamlParser(path).get_config() -> dict``.  Uses ruamel.yaml when present, else PyYAML."""


def _load_all(stream):
    try:
        from ruamel.yaml import YAML
        return list(YAML().load_all(stream))
    except ImportError:
        import yaml
        return list(yaml.safe_load_all(stream))


def _plain(obj):
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    return obj


class YamlParser:
    def __init__(self, path):
        with open(path, "r") as stream:
            docs = [d for d in _load_all(stream) if d is not None]
        # like the reference, the last document of the file wins
        self._config = _plain(docs[-1]) if docs else {}

    def get_config(self):
        return self._config
