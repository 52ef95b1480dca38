'''
Generates ``data_golden.json``: `.data` files written by the REFERENCE codec
(``IntegerEncodedEventSequence.to_file``, composer/dataset/sequence.py:1500-1526,
imported from /root/reference under the shims of ``reference_shims.py``) for
seeded random note sequences, stored as hex, together with what the
reference's own readers return for them (``event_ids_from_file`` :1642-1730 and
``from_file`` :1563-1587).  The fixture travels to machines where the
reference is absent; ``tests/test_pipeline.py`` compares this repository's
codec with it byte for byte.

    python tests/golden/make_data_golden.py
'''

import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_tokenizer_golden import random_note_sequence  # noqa: E402
from reference_shims import load_reference_sequence  # noqa: E402


def main():
    ref = load_reference_sequence()
    if ref is None:
        raise SystemExit('the reference tree is not available')
    rng = random.Random(20261017)
    settings = [(10, 100, 32), (10, 100, 4), (20, 50, 16)]
    cases = []
    with tempfile.TemporaryDirectory() as directory:
        for index in range(12):
            increment, max_steps, bins = settings[index % len(settings)]
            count = 0 if index == 11 else rng.randrange(1, 40)          # the last case is an empty sequence
            sequence = random_note_sequence(ref, rng, count, index % 2 == 0)
            events = sequence.to_event_sequence(increment, max_steps, bins)
            encoded = events.to_integer_encoding()
            path = os.path.join(directory, 'case_%d.data' % index)
            encoded.to_file(path)
            with open(path, 'rb') as handle:
                blob = handle.read()
            ids = [int(v) for v in ref.IntegerEncodedEventSequence.event_ids_from_file(path)[0]]
            back = ref.IntegerEncodedEventSequence.from_file(path, decode=False)
            cases.append({
                'time_step_increment': increment, 'max_time_steps': max_steps, 'velocity_bins': bins,
                'notes': [[n.start, n.end, n.pitch, n.velocity] for n in sequence.notes],
                'sustain_periods': [[s.start, s.end] for s in sequence.sustain_periods],
                'file_hex': blob.hex(),
                'ids': ids,
                'pairs': [[int(a), int(b)] for a, b in back.events],
            })
    with open(os.path.join(HERE, 'data_golden.json'), 'w') as handle:
        json.dump({'generator': 'tests/golden/make_data_golden.py', 'byteorder': sys.byteorder, 'cases': cases}, handle)
    print('wrote %d cases' % len(cases))


if __name__ == '__main__':
    main()
