'''
Generates ``tokenizer_golden.json`` by running the REFERENCE tokenizer
(composer/dataset/sequence.py, imported from /root/reference under the shims of
``reference_shims.py``) on seeded random note sequences: events, ids, the
round trip back to notes, for both sustain-pedal modes and several vocabulary
settings.  The fixture travels to machines where the reference is absent.

    python tests/golden/make_tokenizer_golden.py
'''

import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from reference_shims import load_reference_sequence  # noqa: E402


def random_note_sequence(ref, rng, count, with_sustain):
    notes = []
    time = 0
    for _ in range(count):
        time += rng.choice([0, 0, 7, 10, 35, 120, 480, 1500, 2600])
        duration = rng.choice([5, 10, 90, 250, 1000, 3100])
        notes.append(ref.Note(time, time + duration, rng.randrange(0, 128), rng.randrange(0, 128)))
    sustains = []
    if with_sustain:
        start = 0
        for _ in range(max(1, count // 4)):
            start += rng.choice([50, 400, 900])
            length = rng.choice([100, 700, 2000])
            sustains.append(ref.SustainPeriod(start, start + length))
            start += length
    return ref.NoteSequence(notes, sustains)


def main():
    ref = load_reference_sequence()
    if ref is None:
        raise SystemExit('the reference tree is not available')
    rng = random.Random(20240607)
    cases = []
    settings = [(10, 100, 32), (10, 100, 4), (20, 50, 16), (5, 200, 8)]
    for index in range(48):
        increment, max_steps, bins = settings[index % len(settings)]
        with_sustain = index % 3 != 0
        mode_name = 'EXTEND' if index % 5 == 4 else 'EVENTS'
        sequence = random_note_sequence(ref, rng, rng.randrange(1, 24), with_sustain)
        mode = getattr(ref.NoteSequence.SustainPeriodEncodeMode, mode_name)
        events = sequence.to_event_sequence(increment, max_steps, bins, mode)
        ids = [ref.IntegerEncodedEventSequence.event_to_id(e.type, e.value, events.event_ranges,
                                                            events.event_value_ranges) for e in events.events]
        back = events.to_note_sequence()
        cases.append({
            'time_step_increment': increment, 'max_time_steps': max_steps, 'velocity_bins': bins,
            'sustain_mode': mode_name,
            'notes': [[n.start, n.end, n.pitch, n.velocity] for n in sequence.notes],
            'sustain_periods': [[s.start, s.end] for s in sequence.sustain_periods],
            'events': [[e.type.name, e.value] for e in events.events],
            'ids': ids,
            'vocab_size': ref.OneHotEncodedEventSequence.get_one_hot_size(events.event_ranges),
            'round_trip_notes': [[n.start, n.end, n.pitch, n.velocity] for n in back.notes],
            'round_trip_sustains': [[s.start, s.end] for s in back.sustain_periods],
        })
    with open(os.path.join(HERE, 'tokenizer_golden.json'), 'w') as handle:
        json.dump({'generator': 'tests/golden/make_tokenizer_golden.py', 'cases': cases}, handle)
    print('wrote %d cases' % len(cases))


if __name__ == '__main__':
    main()
