"""Phone-label helpers with the reference's names (transcription2phonemes.py:7-27): the dictionary file lists the
pronunciations of the GRID vocabulary; the label ids 0..32 are the indices of the sorted set of its phoneme symbols
(blank = 33 is appended by check_trainconfiguration).  Host-side text utilities used by the inference drivers to
write `.lbl` transcriptions (inference_siasr_ctc.py:246-259)."""
import numpy as np


def load_dictionary(filename):
    """Sorted list of the distinct whitespace-separated symbols of the file (transcription2phonemes.py:7-14)."""
    with open(filename, 'r') as f:
        dictionary = f.read()
    phonemes = dictionary.replace('\n', ' ').split(' ')
    return [ph for ph in sorted(set(phonemes)) if ph != '']


def get_labels(phonemes, dictionary):
    """'B,IH,N,SP,...' -> label ids; the short-pause symbol SP is dropped (transcription2phonemes.py:17-22)."""
    labels = [lab for lab in phonemes.replace('SP', '').split(',') if lab != '']
    return np.asarray([dictionary.index(ph) for ph in labels])


def get_phonemes_from_labels(labels, dictionary):
    return [dictionary[int(x)] for x in labels]
