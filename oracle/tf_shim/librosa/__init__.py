"""The closed-form librosa helpers the reference's core_test.py compares against
(librosa.midi_to_hz / hz_to_midi / amplitude_to_db, restated from librosa's
documented formulae); everything else is absent."""
import numpy as np


def midi_to_hz(notes):
  return 440.0 * (2.0 ** ((np.asanyarray(notes) - 69.0) / 12.0))


def hz_to_midi(frequencies):
  with np.errstate(divide='ignore'):
    return 12 * (np.log2(np.asanyarray(frequencies)) - np.log2(440.0)) + 69


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
  """librosa.amplitude_to_db = power_to_db(S**2, ref**2, amin**2, top_db)."""
  power = np.square(np.abs(np.asarray(S)))
  log_spec = 10.0 * np.log10(np.maximum(amin**2, power))
  log_spec -= 10.0 * np.log10(np.maximum(amin**2, ref**2))
  if top_db is not None:
    log_spec = np.maximum(log_spec, log_spec.max() - top_db)
  return log_spec


def db_to_amplitude(S_db, ref=1.0):
  """librosa.db_to_amplitude = db_to_power(S_db, ref**2) ** 0.5."""
  return (ref**2 * np.power(10.0, 0.1 * np.asarray(S_db)))**0.5


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
  log_spec = 10.0 * np.log10(np.maximum(amin, np.asarray(S)))
  log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
  if top_db is not None:
    log_spec = np.maximum(log_spec, log_spec.max() - top_db)
  return log_spec


def db_to_power(S_db, ref=1.0):
  return ref * np.power(10.0, 0.1 * np.asarray(S_db))


def __getattr__(name):
  raise NotImplementedError('librosa.%s is not available (tf_shim stub)' % name)
