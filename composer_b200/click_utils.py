'''
A click parameter type that parses the name of an ``Enum`` member.

Same contract as the reference's ``EnumType`` (composer/click_utils.py:10-82):
choices are the member *names* (lower-cased when case-insensitive), the
converted value is the member itself, and the metavar is the snake-cased,
upper-cased class name (``ModelType`` -> ``MODEL_TYPE``).
'''

import re
from enum import EnumMeta

import click


class EnumType(click.Choice):
    def __init__(self, enum, casesensitive=True):
        if not isinstance(enum, EnumMeta):
            raise TypeError('`enum` must be an Enum class')

        self.enum = enum
        self.casesensitive = casesensitive
        names = [member.name for member in enum]
        if not casesensitive:
            names = [name.lower() for name in names]

        super().__init__(sorted(set(names)))

    def convert(self, value, param, ctx):
        if isinstance(value, self.enum):
            return value

        if not self.casesensitive:
            value = value.lower()

        chosen = super().convert(value, param, ctx)
        for member in self.enum:
            name = member.name if self.casesensitive else member.name.lower()
            if name == chosen:
                return member

        self.fail('{!r} is not a member of {}'.format(value, self.enum.__name__), param, ctx)

    def get_metavar(self, param, ctx=None):
        words = re.sub(r'([A-Z]+)([A-Z][a-z])', r'\1_\2', self.enum.__name__)
        words = re.sub(r'([a-z\d])([A-Z])', r'\1_\2', words)
        parts = words.replace('-', '_').lower().split('_')
        if parts[-1] == 'enum':
            parts.pop()

        return '_'.join(parts).upper()
