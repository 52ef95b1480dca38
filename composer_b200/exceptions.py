'''
Exception types raised at the CLI / dataset boundary.

Same two names as the reference (composer/exceptions.py:6-20) so callers that
catch them keep working.
'''


class InvalidParameterError(Exception):
    '''An argument had a value the callee cannot work with.'''


class DatasetError(Exception):
    '''A dataset directory or file is missing or malformed.'''
