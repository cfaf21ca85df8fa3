import importlib

registry = {}


def register(id, entry_point=None, max_episode_steps=None, **kwargs):
    registry[id] = dict(entry_point=entry_point,
                        max_episode_steps=max_episode_steps, kwargs=kwargs)


def make(id, **kwargs):
    spec = registry[id]
    mod_name, attr = spec["entry_point"].split(":")
    cls = getattr(importlib.import_module(mod_name), attr)
    return cls(**{**spec["kwargs"], **kwargs})
