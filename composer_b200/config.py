'''
YAML configuration loading with attribute access.

Mirrors the reference's ``composer.config`` interface (composer/config.py:8-72):
``get(path)`` returns a ``ConfigInstance`` whose nested mappings are reachable
both as ``cfg['a']['b']`` and ``cfg.a.b``, and which remembers ``filepath``.
'''

import yaml


class Dotdict(dict):
    '''A ``dict`` whose keys are also attributes (recursively).'''

    def __init__(self, data=None):
        super().__init__()
        for key, value in (data or {}).items():
            self[key] = Dotdict(value) if hasattr(value, 'keys') else value

    def __getattr__(self, name):
        # Reference behaviour: a missing key raises KeyError, not AttributeError.
        return self[name]

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


class ConfigInstance(Dotdict):
    '''A loaded configuration file; ``filepath`` is the file it came from.'''

    def __init__(self, filepath, data):
        super().__init__(data)
        self['filepath'] = filepath


def get(filepath):
    '''
    Loads ``filepath`` (one or more YAML documents, merged top-level key by
    key in order) and returns a :class:`ConfigInstance`.
    '''

    merged = {}
    with open(filepath) as handle:
        for document in yaml.safe_load_all(handle):
            if document:
                merged.update(document)

    return ConfigInstance(filepath, merged)
