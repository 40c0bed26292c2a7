"""Helpers of the reference's local/utils.py that the hot-path entry points use (benchmark decorator, config
checks, audio range squeeze, channel selection).  LSL / DTW helpers are out of scope."""
import functools
import logging
import os
import re
import struct
import sys
import time
from contextlib import contextmanager

import numpy as np

logger = logging.getLogger('utils.py')


def in_offline_mode(config):
    if not config.has_option('Development', 'seeg_file'):
        return False
    if not os.path.exists(config['Development']['seeg_file']):
        print('WARNING: File path to the sEEG file is invalid. Please specify a proper path. Script will exit!')
        exit(1)
    return True


def select_channels(ch_names, good_channels):
    patterns = [re.compile(r'^{}$'.format(gc)) for gc in good_channels]
    return [name for name in ch_names if any(p.match(name) for p in patterns)]


def squeeze_audio_to_float64(audio):
    if audio.dtype.kind == 'i':
        audio = audio / (2 ** 15)
        logger.info('Input audio has an integer encoding. Converted to float32.')
    if np.max(audio) > 1:
        logger.warning('Expecting audio to be in the range (-1, 1). However, the maximum value is {}'.format(np.max(audio)))
        audio = audio / (2 ** 15)
    if np.min(audio) < -1:
        logger.warning('Expecting audio to be in the range (-1, 1). However, the minimum value is {}.'.format(np.min(audio)))
        audio = audio / (2 ** 15)
    return audio


def check_if_python_shell_is_x64():
    mode = struct.calcsize("P") * 8
    if mode != 64:
        logger.warning('Python Shell is running in x{} and not in x64.'.format(mode))
    else:
        logger.info('Python Shell running in x{} (as recommended)'.format(mode))


@contextmanager
def suppress_stdout():
    with open(os.devnull, 'w') as devnull:
        stdout = sys.stdout
        sys.stdout = devnull
        try:
            yield
        finally:
            sys.stdout = stdout


def benchmark(func):
    """Logs the wall time of the decorated function (reference: local/utils.py:108-121)."""
    @functools.wraps(func)
    def timed(*args, **kwargs):
        t0 = time.perf_counter()
        value = func(*args, **kwargs)
        logger.info('Finished method [{}] in {:.4f} seconds.'.format(func.__name__, time.perf_counter() - t0))
        return value
    return timed
