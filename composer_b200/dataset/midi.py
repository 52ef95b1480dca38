'''
Minimal standard-MIDI-file reader / writer for the two places the hot path
touches MIDI: the prompt of ``composer generate`` (reference:
``NoteSequence.from_midi``, composer/dataset/sequence.py:627-680) and its output
(``NoteSequence.to_midi``, :594-625).  The reference delegates to ``pretty_midi``,
which is not installable here; this module parses / emits SMF directly
(formats 0 and 1, tempo map, running status) with times in milliseconds.
'''

import struct

from composer_b200.dataset import sequence

PPQ = 480
DEFAULT_TEMPO = 500000      # microseconds per quarter note (120 bpm), pretty_midi's default as well


def _vlq(value):
    value = int(value)
    out = [value & 0x7F]
    value >>= 7
    while value:
        out.append((value & 0x7F) | 0x80)
        value >>= 7
    return bytes(reversed(out))


def write_note_sequence(note_sequence, filepath, program=1):
    '''One track: program change, note on/off, CC 64 (64 = pedal down, 63 = up) like the reference writes them.'''

    ticks_per_ms = PPQ * 1000.0 / DEFAULT_TEMPO
    events = []     # (tick, order, bytes)
    for note in note_sequence.notes:
        pitch = int(min(max(note.pitch, 0), 127))
        velocity = int(min(max(note.velocity, 0), 127))
        events.append((int(round(note.start * ticks_per_ms)), 2, bytes([0x90, pitch, velocity])))
        events.append((int(round(note.end * ticks_per_ms)), 0, bytes([0x80, pitch, 0])))
    for period in note_sequence.sustain_periods:
        events.append((int(round(period.start * ticks_per_ms)), 1, bytes([0xB0, 64, 64])))
        events.append((int(round(period.end * ticks_per_ms)), 1, bytes([0xB0, 64, 63])))
    events.sort(key=lambda item: (item[0], item[1]))

    track = bytearray()
    track += _vlq(0) + b'\xFF\x51\x03' + struct.pack('>I', DEFAULT_TEMPO)[1:]
    track += _vlq(0) + bytes([0xC0, int(program) & 0x7F])
    cursor = 0
    for tick, _, payload in events:
        track += _vlq(max(tick - cursor, 0)) + payload
        cursor = max(cursor, tick)
    track += _vlq(0) + b'\xFF\x2F\x00'
    with open(filepath, 'wb') as handle:
        handle.write(b'MThd' + struct.pack('>IHHH', 6, 0, 1, PPQ))
        handle.write(b'MTrk' + struct.pack('>I', len(track)) + bytes(track))


def _read_vlq(data, pos):
    value = 0
    while True:
        byte = data[pos]
        pos += 1
        value = (value << 7) | (byte & 0x7F)
        if not byte & 0x80:
            return value, pos


def _parse_track(data):
    '''Yields (absolute_tick, status, data1, data2 | payload) for channel and tempo events.'''

    pos, tick, status = 0, 0, None
    out = []
    while pos < len(data):
        delta, pos = _read_vlq(data, pos)
        tick += delta
        byte = data[pos]
        if byte == 0xFF:
            kind = data[pos + 1]
            length, pos = _read_vlq(data, pos + 2)
            payload = data[pos:pos + length]
            pos += length
            if kind == 0x51 and length == 3:
                out.append((tick, 0xFF51, int.from_bytes(payload, 'big'), 0))
            if kind == 0x2F:
                break
            continue
        if byte in (0xF0, 0xF7):
            length, pos = _read_vlq(data, pos + 1)
            pos += length
            continue
        if byte & 0x80:
            status = byte
            pos += 1
        if status is None:
            raise ValueError('malformed MIDI track: data byte without a status')
        kind = status & 0xF0
        if kind in (0xC0, 0xD0):
            out.append((tick, status, data[pos], 0))
            pos += 1
        else:
            out.append((tick, status, data[pos], data[pos + 1]))
            pos += 2
    return out


def read_note_sequence(filepath, programs=None, ignore_drums=True):
    with open(filepath, 'rb') as handle:
        data = handle.read()
    if data[:4] != b'MThd':
        raise ValueError('\'{}\' is not a standard MIDI file'.format(filepath))
    header_length, _, track_count, division = struct.unpack('>IHHH', data[4:14])
    if division & 0x8000:
        raise ValueError('SMPTE time division is not supported')
    pos = 8 + header_length
    tracks = []
    for _ in range(track_count):
        if data[pos:pos + 4] != b'MTrk':
            break
        length = struct.unpack('>I', data[pos + 4:pos + 8])[0]
        tracks.append(_parse_track(data[pos + 8:pos + 8 + length]))
        pos += 8 + length

    # tempo map over all tracks -> tick to milliseconds
    tempo_changes = sorted((tick, value) for track in tracks for tick, status, value, _ in track if status == 0xFF51)
    segments = [(0, 0.0, DEFAULT_TEMPO)]       # (start tick, start ms, tempo)
    for tick, tempo in tempo_changes:
        start_tick, start_ms, current = segments[-1]
        elapsed = start_ms + (tick - start_tick) * current / division / 1000.0
        if tick == start_tick:
            segments[-1] = (tick, start_ms, tempo)
        else:
            segments.append((tick, elapsed, tempo))

    def to_ms(tick):
        for start_tick, start_ms, tempo in reversed(segments):
            if tick >= start_tick:
                return start_ms + (tick - start_tick) * tempo / division / 1000.0
        return 0.0

    notes, sustains = [], []
    for track in tracks:
        program_of = {}
        active = {}
        pedal = {}
        for tick, status, a, b in track:
            if status == 0xFF51:
                continue
            kind, channel = status & 0xF0, status & 0x0F
            if kind == 0xC0:
                program_of[channel] = a
                continue
            if ignore_drums and channel == 9:
                continue
            if programs is not None and program_of.get(channel, 0) not in programs:
                continue
            now = to_ms(tick)
            if kind == 0x90 and b > 0:
                active.setdefault((channel, a), []).append((now, b))
            elif kind == 0x80 or (kind == 0x90 and b == 0):
                pending = active.get((channel, a))
                if pending:
                    start, velocity = pending.pop(0)
                    notes.append(sequence.Note(start, now, a, velocity))
            elif kind == 0xB0 and a == 64:
                # same rule as the reference (sequence.py:659-678)
                current = pedal.get(channel)
                if b >= 64 and current is None:
                    pedal[channel] = sequence.SustainPeriod(now, None)
                elif b < 64:
                    if current is not None:
                        current.end = now
                        sustains.append(current)
                        pedal[channel] = None
                    elif sustains:
                        sustains[-1].end = now
    notes.sort(key=lambda note: note.start)
    return sequence.NoteSequence(notes, sustains)
