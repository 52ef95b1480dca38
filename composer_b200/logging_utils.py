'''
Root-logger formatting for the CLI.

Behaviour follows the reference (composer/logging_utils.py:26-52): INFO records
print the bare message, other levels are prefixed with a coloured level name.
``colorama`` is optional here; plain ANSI codes are used when it is absent.
'''

import copy
import logging

try:  # pragma: no cover - depends on the environment
    import colorama
    _RESET = colorama.Style.RESET_ALL
    _COLOURS = {
        logging.FATAL: colorama.Fore.LIGHTRED_EX,
        logging.ERROR: colorama.Fore.RED,
        logging.WARNING: colorama.Fore.YELLOW,
        logging.DEBUG: colorama.Fore.LIGHTWHITE_EX,
    }
    GREEN = colorama.Fore.GREEN
except ImportError:
    _RESET = '\033[0m'
    _COLOURS = {
        logging.FATAL: '\033[91m',
        logging.ERROR: '\033[31m',
        logging.WARNING: '\033[33m',
        logging.DEBUG: '\033[97m',
    }
    GREEN = '\033[32m'

_DEFAULT_FORMAT = '%(levelname)s: %(message)s'
_PER_LEVEL_FORMAT = {logging.INFO: '%(message)s'}


def colourize_string(string, colour):
    return '{}{}{}'.format(colour, string, _RESET)


class _LevelAwareFormatter(logging.Formatter):
    def format(self, record):
        record = copy.copy(record)
        colour = _COLOURS.get(record.levelno)
        if colour is not None:
            record.levelname = colourize_string(record.levelname, colour)

        saved = self._style._fmt
        self._style._fmt = _PER_LEVEL_FORMAT.get(record.levelno, saved)
        try:
            return super().format(record)
        finally:
            self._style._fmt = saved


_installed = False


def init():
    '''Installs the stream handler on the root logger (idempotent).'''

    global _installed
    if _installed:
        return

    handler = logging.StreamHandler()
    handler.setFormatter(_LevelAwareFormatter(_DEFAULT_FORMAT))
    logging.getLogger().addHandler(handler)
    _installed = True
