"""Mirror of the reference plug-in point ``src/models/ModelFactory.py:10-22``: the model file is chosen by
``configs['model']['name']`` and the class is that name minus its two-character version suffix."""
import importlib
import inspect


def get_model(configs: dict, model_configs: dict = None):
    filename = configs['model']['name']
    classname = filename[:-2]
    try:
        module = importlib.import_module(f'{__package__}.{filename}')
    except ModuleNotFoundError as exc:
        raise RuntimeError(f'Unknown model: {filename}') from exc
    for name, cls in inspect.getmembers(module, inspect.isclass):
        if name == classname:
            return cls(configs, model_configs)
    raise RuntimeError(f'Unknown model: {filename}')
