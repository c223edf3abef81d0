"""Config objects with the reference's semantics (MFT/config.py:7-52 of serycjon/MFT):
a missing attribute reads as an empty, falsy Config, so optional flags default to "off";
load_config executes a python file and returns its get_config()."""
import importlib.util
from pathlib import Path


class Config:
    def __getattr__(self, name):
        # only reached for attributes that were never set
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return Config()

    def __bool__(self):
        return False

    def __repr__(self):
        return repr(self.__dict__)

    def __eq__(self, other):
        return isinstance(other, Config) and self.__dict__ == other.__dict__

    def merge(self, other, update_dicts=False):
        for key, value in other.__dict__.items():
            mine = self.__dict__.get(key)
            if update_dicts and isinstance(mine, dict) and isinstance(value, dict):
                mine.update(value)
            else:
                setattr(self, key, value)


def load_config(path):
    path = Path(path)
    if not path.exists():
        raise AssertionError(f'config {path} does not exist!')
    spec = importlib.util.spec_from_file_location('tracker_config', str(path))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module.get_config()
