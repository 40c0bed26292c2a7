"""The one helper of the reference's local/utils.py that the hot-path entry points use: the @benchmark stage timer
(reference local/utils.py:108-121, applied at train.py:27-120 and decode.py:32).  Everything else in that file (LSL / DTW
helpers, config checks, channel-name matching) is out of scope and is not provided."""
import functools
import logging
import time

logger = logging.getLogger('utils.py')


def benchmark(func):
    """Logs the wall time of the decorated function."""
    @functools.wraps(func)
    def timed(*args, **kwargs):
        t0 = time.perf_counter()
        value = func(*args, **kwargs)
        logger.info('Finished method [{}] in {:.4f} seconds.'.format(func.__name__, time.perf_counter() - t0))
        return value
    return timed
