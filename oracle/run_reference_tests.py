"""Runs the reference's OWN unit tests for the decoder path (ddsp/core_test.py,
synths_test.py, processors_test.py, dags-related tests) against oracle/tf_shim.

    python -m oracle.run_reference_tests [-v] [module ...]

TEST INFRASTRUCTURE ONLY; needs /root/reference (authoring container).  Prints a
summary and exits non-zero on failures.  Tests that need pieces outside the path
(crepe, librosa beyond the two closed forms, tfp) are expected to error and are
listed separately.
"""
import importlib
import sys
import unittest

from oracle import ref_on_shim

DEFAULT_MODULES = ('ddsp.core_test', 'ddsp.synths_test', 'ddsp.processors_test')


def run(modules=DEFAULT_MODULES, verbosity=1, stream=None):
  ref_on_shim.load()
  suite = unittest.TestSuite()
  loader = unittest.TestLoader()
  for name in modules:
    suite.addTests(loader.loadTestsFromModule(importlib.import_module(name)))
  runner = unittest.TextTestRunner(verbosity=verbosity, stream=stream or sys.stderr)
  return runner.run(suite)


if __name__ == '__main__':
  args = [a for a in sys.argv[1:] if not a.startswith('-')]
  res = run(tuple(args) or DEFAULT_MODULES, verbosity=2 if '-v' in sys.argv else 1)
  print('ran %d, failures %d, errors %d, skipped %d' % (
      res.testsRun, len(res.failures), len(res.errors), len(res.skipped)))
  sys.exit(0 if res.wasSuccessful() else 1)
