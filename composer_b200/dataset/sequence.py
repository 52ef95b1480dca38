'''
Note sequences, event sequences and the integer event vocabulary.

This is the tokenizer side of the Transformer hot path: it defines which
integer id every (event type, value) pair gets, so it has to agree with the
reference bit for bit.  The public names and call signatures follow the
reference module (composer/dataset/sequence.py) so code written against it
keeps working; the implementation is organised around one object,
:class:`EventVocabulary`, that owns the id layout and does encode / decode as
table lookups (scalar or vectorised).

Vocabulary layout (reference: sequence.py:740-766, 792-805, 826-844), in order:

    NOTE_ON      ids [0, 128)                       value = MIDI pitch
    NOTE_OFF     ids [128, 256)                     value = MIDI pitch
    VELOCITY     ids [256, 256 + bins)              value = velocity bin
    TIME_SHIFT   ids [256 + bins, 256 + bins + S)   value = steps 1..S
    SUSTAIN_ON   one id                             value = None
    SUSTAIN_OFF  one id                             value = None

With the default dataset configuration (bins = 32, S = 100) that is 390 ids.
'''

import array
import collections
import copy
import os
import struct
from enum import Enum, IntEnum, unique
from pathlib import Path

import numpy as np

from composer_b200.exceptions import InvalidParameterError


class EventType(IntEnum):
    '''The kind of an :class:`Event` (integer values are part of the file format).'''

    NOTE_ON = 1
    NOTE_OFF = 2
    TIME_SHIFT = 3
    VELOCITY = 4
    SUSTAIN_ON = 5
    SUSTAIN_OFF = 6

    @staticmethod
    def make_int_type_map():
        return {int(member): member for member in EventType}


_EVENT_TYPE_MAPPINGS = EventType.make_int_type_map()


class Note:
    '''A note: ``start`` / ``end`` in milliseconds, MIDI ``pitch`` and ``velocity``.'''

    __slots__ = ('start', 'end', 'pitch', 'velocity')

    def __init__(self, start, end, pitch, velocity):
        self.start = start
        self.end = end
        self.pitch = pitch
        self.velocity = velocity

    @property
    def duration(self):
        return self.end - self.start

    def __repr__(self):
        return 'Note(start={:f}, end={:f}, pitch={}, velocity={})'.format(
            self.start, self.end, self.pitch, self.velocity)


class SustainPeriod:
    '''An interval, in milliseconds, during which the sustain pedal is down.'''

    __slots__ = ('start', 'end')

    def __init__(self, start, end):
        self.start = start
        self.end = end

    def __repr__(self):
        return 'SustainPeriod(start={}, end={})'.format(self.start, self.end)


class Event:
    '''One token of the event language: a type and an (optional) integer value.'''

    NONE_VALUE = -1
    __slots__ = ('type', 'value')

    def __init__(self, event_type, value):
        self.type = event_type
        self.value = value

    @staticmethod
    def encode_value(event):
        return Event.NONE_VALUE if event.value is None else int(event.value)

    @staticmethod
    def decode_value(value):
        return None if value == Event.NONE_VALUE else value

    def __eq__(self, other):
        return isinstance(other, Event) and self.type == other.type and self.value == other.value

    def __hash__(self):
        return hash((int(self.type), self.value))

    def __str__(self):
        return '{}<{}>'.format(self.type.name, self.value)

    def __repr__(self):
        return 'Event(type={}, value={})'.format(str(self.type), self.value)


class EventVocabulary:
    '''
    The id layout for one dataset configuration.

    ``value_ranges`` / ``dimensions`` / ``ranges`` are the three ordered
    mappings the reference computes (sequence.py:740-766, 792-805, 826-844);
    the lookup tables make id <-> (type, value) conversion O(1) and usable on
    whole numpy arrays.
    '''

    _ORDER = (EventType.NOTE_ON, EventType.NOTE_OFF, EventType.VELOCITY,
              EventType.TIME_SHIFT, EventType.SUSTAIN_ON, EventType.SUSTAIN_OFF)

    def __init__(self, time_step_increment=10, max_time_steps=100, velocity_bins=32):
        self.time_step_increment = time_step_increment
        self.max_time_steps = max_time_steps
        self.velocity_bins = velocity_bins

        self.value_ranges = collections.OrderedDict([
            (EventType.NOTE_ON, range(0, 128)),
            (EventType.NOTE_OFF, range(0, 128)),
            (EventType.VELOCITY, range(0, velocity_bins)),
            # A shift of zero steps is never emitted, so values start at 1.
            (EventType.TIME_SHIFT, range(1, max_time_steps + 1)),
            (EventType.SUSTAIN_ON, None),
            (EventType.SUSTAIN_OFF, None),
        ])

        self.dimensions = collections.OrderedDict(
            (kind, 0 if values is None else values.stop - values.start)
            for kind, values in self.value_ranges.items())

        self.ranges = collections.OrderedDict()
        cursor = 0
        for kind, width in self.dimensions.items():
            width = max(width, 1)  # value-less events still occupy one id
            self.ranges[kind] = range(cursor, cursor + width)
            cursor += width

        self.size = cursor

        # id -> type / value tables, and type -> (first id, first value).
        self._type_of_id = np.empty(self.size, dtype=np.int16)
        self._value_of_id = np.full(self.size, Event.NONE_VALUE, dtype=np.int16)
        self._first_id = np.zeros(max(EventType) + 1, dtype=np.int32)
        self._first_value = np.zeros(max(EventType) + 1, dtype=np.int32)
        self._has_value = np.zeros(max(EventType) + 1, dtype=bool)
        for kind, ids in self.ranges.items():
            self._type_of_id[ids.start:ids.stop] = int(kind)
            self._first_id[kind] = ids.start
            values = self.value_ranges[kind]
            if values is not None:
                self._value_of_id[ids.start:ids.stop] = np.arange(values.start, values.stop)
                self._first_value[kind] = values.start
                self._has_value[kind] = True

    def event_to_id(self, event_type, event_value):
        first = self.ranges[event_type].start
        values = self.value_ranges[event_type]
        return first if values is None else first + (event_value - values.start)

    def id_to_event(self, event_id):
        if not 0 <= event_id < self.size:
            return None

        kind = _EVENT_TYPE_MAPPINGS[int(self._type_of_id[event_id])]
        value = None if self.value_ranges[kind] is None else int(self._value_of_id[event_id])
        return Event(kind, value)

    def encode_array(self, types, values):
        '''Vectorised ``event_to_id`` over parallel integer arrays.'''

        types = np.asarray(types, dtype=np.int64)
        values = np.asarray(values, dtype=np.int64)
        offsets = np.where(self._has_value[types], values - self._first_value[types], 0)
        return self._first_id[types] + offsets

    def decode_array(self, ids):
        '''Vectorised ``id_to_event``: returns (types, values) int16 arrays; value -1 = None.'''

        ids = np.asarray(ids, dtype=np.int64)
        return self._type_of_id[ids], self._value_of_id[ids]


class NoteSequence:
    '''Notes plus sustain-pedal periods; the MIDI-facing representation.'''

    @unique
    class SustainPeriodEncodeMode(Enum):
        NONE = 'none'
        EXTEND = 'extend'
        EVENTS = 'events'

    def __init__(self, notes=None, sustain_periods=None):
        self.notes = []
        if notes is not None:
            self.add_notes(notes, maintain_order=False)
            self.notes.sort(key=lambda note: note.start)

        self.sustain_periods = sustain_periods if sustain_periods is not None else []

    def add_notes(self, notes, maintain_order=True):
        self.notes.extend(notes)
        if maintain_order:
            self.notes.sort(key=lambda note: note.start)

    def _target(self, inplace, copy_sustains=True):
        if inplace:
            return self.notes, self.sustain_periods

        sustains = copy.deepcopy(self.sustain_periods) if copy_sustains else self.sustain_periods
        return copy.deepcopy(self.notes), sustains

    def time_stretch(self, percent, inplace=True):
        notes, sustains = self._target(inplace)
        for item in (*notes, *sustains):
            item.start *= percent
            item.end *= percent

        return self if inplace else NoteSequence(notes, sustains)

    def time_shift(self, offset, inplace=True):
        notes, sustains = self._target(inplace)
        for item in (*notes, *sustains):
            item.start += offset
            item.end += offset

        return self if inplace else NoteSequence(notes, sustains)

    def trim_start(self, inplace=True):
        '''Shifts everything so the first note (or pedal press) starts at 0.'''

        first = self.notes[0].start
        if self.sustain_periods:
            first = min(first, self.sustain_periods[0].start)

        return self.time_shift(-first, inplace=inplace)

    def pitch_shift(self, offset, inplace=True):
        '''Adds ``offset`` to every pitch, clamped to the MIDI range [0, 127].'''

        notes, sustains = self._target(inplace)
        for note in notes:
            note.pitch = min(max(note.pitch + offset, 0), 127)

        return self if inplace else NoteSequence(notes, sustains)

    def _extend_notes_over_sustain(self, ordered_notes, ordered_sustains):
        # Reference semantics (sequence.py:491-514), including its scan cursor:
        # the cursor only advances when a period contained notes, and then to
        # the index the scan stopped at.
        cursor = 0
        for period in ordered_sustains:
            inside = []
            index = cursor
            for index in range(cursor, len(ordered_notes)):
                note = ordered_notes[index]
                if note.start < period.start:
                    continue
                if note.start > period.end:
                    break
                inside.append(note)

            if not inside:
                continue

            cursor = index
            next_start_of_pitch = {}
            for note in reversed(inside):
                if note.pitch in next_start_of_pitch:
                    note.end = next_start_of_pitch[note.pitch]
                else:
                    note.end = max(period.end, note.end)
                next_start_of_pitch[note.pitch] = note.start

    def to_event_sequence(self, time_step_increment=10, max_time_steps=100, velocity_bins=32,
                          sustain_period_encode_mode=None, clean=True):
        '''
        Flattens notes (and, depending on the mode, pedal periods) into a
        time-ordered event list (reference: sequence.py:383-592).
        '''

        mode = sustain_period_encode_mode
        if mode is None:
            mode = NoteSequence.SustainPeriodEncodeMode.EVENTS

        ordered_notes = sorted(self.notes, key=lambda note: note.start)
        ordered_sustains = sorted(self.sustain_periods, key=lambda period: period.start)

        # (time, kind, on?, payload) markers; Python's sort is stable, so equal
        # times keep insertion order: pedal markers first, then notes in start
        # order with each note's ON directly before its OFF.
        markers = []
        if mode == NoteSequence.SustainPeriodEncodeMode.EVENTS:
            for period in ordered_sustains:
                markers.append((period.start, 'SUSTAIN', True, period))
                markers.append((period.end, 'SUSTAIN', False, period))
        elif mode == NoteSequence.SustainPeriodEncodeMode.EXTEND:
            self._extend_notes_over_sustain(ordered_notes, ordered_sustains)

        for note in ordered_notes:
            markers.append((note.start, 'NOTE', True, note))
            markers.append((note.end, 'NOTE', False, note))

        markers.sort(key=lambda marker: marker[0])

        events = []
        now = 0
        velocity = 0
        for time, kind, is_on, payload in markers:
            # The rounding happens before the division, exactly as the
            # reference does it (sequence.py:530).
            steps = int(round(time - now) / time_step_increment)
            if max_time_steps is not None:
                events.extend(Event(EventType.TIME_SHIFT, max_time_steps)
                              for _ in range(steps // max_time_steps))
                steps %= max_time_steps

            if steps > 0:
                events.append(Event(EventType.TIME_SHIFT, steps))

            if kind == 'NOTE':
                if velocity != payload.velocity:
                    events.append(Event(EventType.VELOCITY, (payload.velocity * velocity_bins) // 128))

                events.append(Event(EventType.NOTE_ON if is_on else EventType.NOTE_OFF, payload.pitch))
                velocity = payload.velocity
            else:
                events.append(Event(EventType.SUSTAIN_ON if is_on else EventType.SUSTAIN_OFF, None))

            now = time

        if clean:
            events = _drop_redundant_events(events)

        return EventSequence(events, time_step_increment, max_time_steps, velocity_bins)

    def to_midi(self, filepath, program=1):
        '''Writes a single-track standard MIDI file (see :mod:`composer_b200.dataset.midi`).'''

        from composer_b200.dataset import midi
        midi.write_note_sequence(self, filepath, program=program)

    @staticmethod
    def from_midi(filepath, programs=None, ignore_drums=True):
        '''Reads notes and pedal (CC 64) periods from a standard MIDI file.'''

        filepath = Path(filepath)
        if not filepath.is_file():
            raise InvalidParameterError(
                'Cannot create NoteSequence from \'{}\' since it is not a file.'.format(filepath))

        from composer_b200.dataset import midi
        return midi.read_note_sequence(filepath, programs=programs, ignore_drums=ignore_drums)


def _drop_redundant_events(events):
    '''
    Removes zero-length time shifts and NOTE_ON/NOTE_OFF (or OFF/ON) pairs of
    the same pitch that sit directly next to each other (sequence.py:566-590).
    Decisions are taken on the *original* list, scanning from the back.

    Bit-exactness note: the reference collects indices in a list, so an event
    that belongs to two overlapping pairs (ON x, OFF x, ON x) is queued twice,
    and each queued index is then popped from the shrinking list in descending
    order.  The second pop of a repeated index therefore removes whatever
    slid into that slot (or raises ``IndexError`` at the end of the list).
    That behaviour is reproduced here, because it decides which ids a dataset
    contains.
    '''

    queued = []
    for i in range(len(events) - 1, -1, -1):
        event = events[i]
        if event.type == EventType.TIME_SHIFT and event.value == 0:
            queued.append(i)

        if i == 0:
            continue

        previous = events[i - 1]
        flipped = (event.type == EventType.NOTE_OFF and previous.type == EventType.NOTE_ON) or \
                  (event.type == EventType.NOTE_ON and previous.type == EventType.NOTE_OFF)
        if flipped and event.value == previous.value:
            queued.extend((i, i - 1))

    kept = list(events)
    for i in sorted(queued, reverse=True):
        kept.pop(i)

    return kept


class EventSequence:
    '''A list of :class:`Event` plus the three settings that define its vocabulary.'''

    def __init__(self, events, time_step_increment, max_time_steps, velocity_bins):
        self.events = events
        self.time_step_increment = time_step_increment
        self.max_time_steps = max_time_steps
        self.velocity_bins = velocity_bins

    # -- vocabulary -------------------------------------------------------

    @staticmethod
    def _compute_event_value_ranges(time_step_increment, max_time_steps, velocity_bins):
        return EventVocabulary(time_step_increment, max_time_steps, velocity_bins).value_ranges

    @staticmethod
    def _compute_event_dimensions(event_value_ranges):
        return collections.OrderedDict(
            (kind, 0 if values is None else values.stop - values.start)
            for kind, values in event_value_ranges.items())

    @staticmethod
    def _compute_event_ranges(event_dimensions):
        ranges = collections.OrderedDict()
        cursor = 0
        for kind, width in event_dimensions.items():
            width = max(width, 1)
            ranges[kind] = range(cursor, cursor + width)
            cursor += width

        return ranges

    @property
    def vocabulary(self):
        limit = self.max_time_steps
        if limit is None:
            limit = max(event.value for event in self.events if event.type == EventType.TIME_SHIFT)

        return EventVocabulary(self.time_step_increment, limit, self.velocity_bins)

    @property
    def event_value_ranges(self):
        return self.vocabulary.value_ranges

    @property
    def event_dimensions(self):
        return self.vocabulary.dimensions

    @property
    def event_ranges(self):
        return self.vocabulary.ranges

    # -- conversions ------------------------------------------------------

    def to_integer_encoding(self):
        return IntegerEncodedEventSequence.encode(self)

    def to_ids(self, dtype=np.int32):
        '''All events as vocabulary ids (numpy array).'''

        vocabulary = self.vocabulary
        return np.fromiter((vocabulary.event_to_id(event.type, event.value) for event in self.events),
                           dtype=dtype, count=len(self.events))

    def to_note_sequence(self):
        '''Replays the events into notes and pedal periods (sequence.py:867-924).'''

        now = 0
        velocity = 0
        sounding = {}
        pedal = None
        notes = []
        sustains = []
        for event in self.events:
            if event.type == EventType.NOTE_ON:
                if sounding.get(event.value) is None:
                    sounding[event.value] = Note(now, 0, event.value, velocity)
            elif event.type == EventType.NOTE_OFF:
                note = sounding.get(event.value)
                if note is not None:
                    note.end = now
                    notes.append(note)
                    sounding[event.value] = None
            elif event.type == EventType.TIME_SHIFT:
                now += event.value * self.time_step_increment
            elif event.type == EventType.VELOCITY:
                velocity = (128 * event.value) // self.velocity_bins
            elif event.type == EventType.SUSTAIN_ON:
                if pedal is None:
                    pedal = SustainPeriod(now, 0)
            elif event.type == EventType.SUSTAIN_OFF:
                if pedal is not None:
                    pedal.end = now
                    sustains.append(pedal)
                    pedal = None

        return NoteSequence(notes, sustains)

    @staticmethod
    def from_file(filepath, decode=True):
        with open(filepath, 'rb') as handle:
            type_id = _read_encoding_type_id(handle)

        if type_id != IntegerEncodedEventSequence.get_encoding_type():
            raise InvalidEncodingTypeError(
                'Cannot load \'{}\' as an EventSequence! \'{}\' is not a supported encoding type id.'
                .format(filepath, type_id))

        return IntegerEncodedEventSequence.from_file(filepath, decode=decode)

    def __repr__(self):
        return '\n'.join(str(event) for event in self.events)


class InvalidEncodingTypeError(Exception):
    '''A serialized sequence starts with an unknown encoding type id.'''


_TYPE_ID_FORMAT = 'Q'


def _read_encoding_type_id(handle):
    size = struct.calcsize(_TYPE_ID_FORMAT)
    return struct.unpack(_TYPE_ID_FORMAT, handle.read(size))[0]


class IntegerEncodedEventSequence:
    '''
    The on-disk form of an :class:`EventSequence`: ``(type, value)`` int16 pairs.

    File layout (native byte order and alignment, as written by
    ``struct.pack('Q' + 'hhh' + 'hh' * n, ...)``; reference: sequence.py:1441-1442,
    1500-1526): an unsigned 64-bit encoding type id, then time_step_increment,
    max_time_steps, velocity_bins as int16, then one int16 pair per event.
    Packing everything in one call means there is no padding after the 8-byte
    id, so the events start at byte 14.
    '''

    _HEADER_FORMAT = 'hhh'
    _EVENT_FORMAT = 'hh'
    _ENCODING_TYPE_ID = 9223372036854775805

    def __init__(self, time_step_increment, max_time_steps, velocity_bins, events=None):
        self.time_step_increment = time_step_increment
        self.max_time_steps = max_time_steps
        self.velocity_bins = velocity_bins
        self.events = events if events is not None else []

    @staticmethod
    def get_encoding_type():
        return IntegerEncodedEventSequence._ENCODING_TYPE_ID

    @staticmethod
    def encode(event_sequence):
        pairs = [(int(event.type), Event.encode_value(event)) for event in event_sequence.events]
        return IntegerEncodedEventSequence(event_sequence.time_step_increment, event_sequence.max_time_steps,
                                           event_sequence.velocity_bins, pairs)

    def decode(self):
        events = [Event(_EVENT_TYPE_MAPPINGS[kind], Event.decode_value(value)) for kind, value in self.events]
        return EventSequence(events, self.time_step_increment, self.max_time_steps, self.velocity_bins)

    def to_file(self, filepath):
        flat = np.asarray(self.events, dtype=np.int16).reshape(-1)
        header = struct.pack(_TYPE_ID_FORMAT + self._HEADER_FORMAT, self.get_encoding_type(),
                             self.time_step_increment, self.max_time_steps, self.velocity_bins)
        with open(filepath, 'wb+') as handle:
            handle.write(header)
            handle.write(flat.tobytes())

    @classmethod
    def _header_size(cls):
        return struct.calcsize(_TYPE_ID_FORMAT) + struct.calcsize(cls._HEADER_FORMAT)

    @classmethod
    def _read_raw(cls, filepath):
        '''Returns (settings, int16 array of shape [n, 2]) for a ``.data`` file.'''

        with open(filepath, 'rb') as handle:
            type_id = _read_encoding_type_id(handle)
            if type_id != cls.get_encoding_type():
                raise InvalidEncodingTypeError(
                    'Cannot decode \'{}\' as IntegerEncodedEventSequence since the encoding type id '
                    'header does not match.'.format(filepath))

            settings = struct.unpack(cls._HEADER_FORMAT, handle.read(struct.calcsize(cls._HEADER_FORMAT)))
            payload = handle.read()

        pair_bytes = struct.calcsize(cls._EVENT_FORMAT)
        count = len(payload) // pair_bytes
        pairs = np.frombuffer(payload, dtype=np.int16, count=count * 2).reshape(count, 2)
        return settings, pairs

    @classmethod
    def from_file(cls, filepath, decode=False):
        settings, pairs = cls._read_raw(filepath)
        encoded = cls(*settings, events=[(int(kind), int(value)) for kind, value in pairs])
        return encoded.decode() if decode else encoded

    @staticmethod
    def event_to_id(event_type, event_value, event_ranges, event_value_ranges):
        '''id = first id of the type + (value - first value of the type) (sequence.py:1589-1612).'''

        values = event_value_ranges[event_type]
        offset = 0 if values is None else event_value - values.start
        return event_ranges[event_type].start + offset

    @staticmethod
    def id_to_event(event_id, event_ranges, event_value_ranges):
        '''Inverse of :meth:`event_to_id`; ``None`` for an id outside every range (sequence.py:1614-1640).'''

        for kind, ids in event_ranges.items():
            if event_id in ids:
                values = event_value_ranges[kind]
                value = None if values is None else event_id - ids.start + values.start
                return Event(kind, value)

        return None

    @classmethod
    def event_ids_from_file(cls, filepath, as_numpy_array=False, numpy_dtype=int):
        '''
        Loads a ``.data`` file straight to vocabulary ids.

        Returns ``(ids, event_value_ranges, event_ranges, settings)`` like the
        reference (sequence.py:1642-1698); ``ids`` is an ``array('H')`` unless
        ``as_numpy_array`` is set.
        '''

        settings, pairs = cls._read_raw(filepath)
        vocabulary = EventVocabulary(*settings)
        ids = vocabulary.encode_array(pairs[:, 0], pairs[:, 1])
        if as_numpy_array:
            ids = ids.astype(numpy_dtype)
        else:
            ids = array.array('H', ids.astype(np.uint16).tobytes())

        return ids, vocabulary.value_ranges, vocabulary.ranges, settings

    @classmethod
    def event_ids_from_file_as_generator(cls, filepath):
        ids, _, _, _ = cls.event_ids_from_file(filepath, as_numpy_array=True, numpy_dtype=np.int64)
        for event_id in ids:
            yield int(event_id)


class OneHotEncodedEventSequence:
    '''Only the vocabulary-size helper of the reference class is needed on this path.'''

    @staticmethod
    def get_one_hot_size(event_ranges):
        '''Vocabulary size = end of the last id range (sequence.py:1120-1130).'''

        return event_ranges[next(reversed(event_ranges))].stop
