'''
Input pipeline for ``composer train transformer <dataset dir>``: preprocessed
``.data`` files -> one stream of event ids -> non-overlapping windows of
``window_size + 1`` ids -> ``(x, y)`` with ``y`` = ``x`` shifted by one ->
shuffle buffer of ``500 * batch_size`` windows -> batches (remainder dropped).

Mirrors the reference's ``load_events`` / ``load_dataset``
(composer/models/__init__.py:160-313) without TensorFlow: numpy on the host,
pinned-memory staging so that the host-to-device copy of a batch overlaps the
previous step.  ``rank`` / ``world_size`` shard the batches for data-parallel
training (every rank sees the same number of batches).
'''

import logging
import threading
import queue
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from composer_b200.dataset.sequence import IntegerEncodedEventSequence
from composer_b200.exceptions import InvalidParameterError

OUTPUT_EXTENSION = 'data'       # composer/dataset/preprocess.py: processed files are ``*.data``


def get_processed_files(dataset_path):
    '''All preprocessed files below ``dataset_path`` (composer/dataset/preprocess.py:18-34).'''

    dataset_path = Path(dataset_path)
    if not dataset_path.is_dir():
        raise InvalidParameterError('\'{}\' is an invalid dataset path!'.format(dataset_path))

    return sorted(dataset_path.glob('**/*.{}'.format(OUTPUT_EXTENSION)))


def load_events(filepaths, show_loading_progress_bar=False, n_jobs=16):
    '''Concatenates the ids of every file, in order, into one uint16 array (models/__init__.py:227-233).'''

    filepaths = list(filepaths)
    if not filepaths:
        return np.zeros(0, dtype=np.uint16)

    def load(path):
        ids, _, _, _ = IntegerEncodedEventSequence.event_ids_from_file(str(path), as_numpy_array=True,
                                                                       numpy_dtype=np.uint16)
        return ids

    logging.info('- Loading dataset (\'{}\') into memory.'.format(Path(filepaths[0]).parent))
    with ThreadPoolExecutor(max_workers=n_jobs) as pool:
        parts = list(pool.map(load, filepaths))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint16)


class EventWindowDataset:
    '''
    Re-iterable dataset of ``(x, y)`` int32 batches ``[batch_size, window_size]``.
    Every ``iter()`` is one epoch; with ``shuffle`` the windows pass through a
    buffer of ``500 * batch_size`` entries that is re-drawn each epoch
    (``reshuffle_each_iteration=True`` in the reference).
    '''

    def __init__(self, events, batch_size, window_size, shuffle=True, seed=0, rank=0, world_size=1,
                 shuffle_buffer_batches=500):
        self.events = np.asarray(events)
        self.batch_size = int(batch_size)
        self.window_size = int(window_size)
        self.shuffle = shuffle
        self.seed = int(seed)
        self.rank, self.world_size = int(rank), int(world_size)
        self.shuffle_buffer = shuffle_buffer_batches * self.batch_size
        self.epoch = 0
        span = self.window_size + 1
        self.window_count = len(self.events) // span                      # batch(window+1, drop_remainder=True)
        self.windows = self.events[:self.window_count * span].reshape(self.window_count, span)

    def __len__(self):
        return (self.window_count // self.batch_size) // self.world_size

    def _order(self):
        order = np.arange(self.window_count)
        if not self.shuffle or self.window_count == 0:
            return order
        # streaming shuffle buffer, as tf.data.Dataset.shuffle does it; all ranks draw the same order
        rng = np.random.default_rng((self.seed, self.epoch))
        buffer = list(order[:self.shuffle_buffer])
        out = []
        for item in order[self.shuffle_buffer:]:
            slot = int(rng.integers(len(buffer)))
            out.append(buffer[slot])
            buffer[slot] = item
        rng.shuffle(buffer)
        out.extend(buffer)
        return np.asarray(out, dtype=np.int64)

    def __iter__(self):
        order = self._order()
        self.epoch += 1
        batches = len(order) // self.batch_size                           # batch(batch_size, drop_remainder=True)
        batches -= batches % self.world_size
        for index in range(self.rank, batches, self.world_size):
            rows = order[index * self.batch_size:(index + 1) * self.batch_size]
            block = self.windows[rows].astype(np.int32)
            yield block[:, :-1], block[:, 1:]


def prefetch_pinned(dataset, depth=2):
    '''
    Wraps a dataset so that batches are staged into pinned host memory by a
    background thread (``depth`` batches ahead); ``train_step`` then issues a
    non-blocking copy.  Re-iterable like the dataset it wraps.
    '''

    import torch

    class _Prefetcher:
        def __len__(self):
            return len(dataset)

        def __iter__(self):
            slots = queue.Queue(maxsize=depth)
            end = object()
            stop = threading.Event()

            def put(item):
                # never blocks for good: the consumer may have stopped (max_steps, an exception) and left
                while not stop.is_set():
                    try:
                        slots.put(item, timeout=0.1)
                        return True
                    except queue.Full:
                        continue
                return False

            def produce():
                try:
                    for x, y in dataset:
                        pair = (torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(y)))
                        if torch.cuda.is_available():
                            pair = (pair[0].pin_memory(), pair[1].pin_memory())
                        if not put(pair):
                            return
                    put(end)
                except BaseException as error:   # a bad file, pinned-memory exhaustion: surfaces in the consumer
                    put(error)

            thread = threading.Thread(target=produce, daemon=True)
            thread.start()
            try:
                while True:
                    item = slots.get()
                    if item is end:
                        return
                    if isinstance(item, BaseException):
                        raise item
                    yield item
            finally:
                stop.set()

    return _Prefetcher()


def load_dataset(filepaths, batch_size, window_size, show_loading_progress_bar=True, shuffle=True, seed=0, rank=0,
                 world_size=1):
    '''The reference's ``load_dataset`` for integer-encoded events (models/__init__.py:238-313).'''

    events = load_events(filepaths, show_loading_progress_bar)
    return EventWindowDataset(events, batch_size, window_size, shuffle=shuffle, seed=seed, rank=rank,
                              world_size=world_size)
