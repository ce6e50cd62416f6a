"""Checks that need the reference sources (/root/reference): they run in the build container's CPU tier and
skip everywhere else (the GPU box has no /root/reference).  Each runs in a subprocess because the reference
package and this repository's alias package share the name ``kapre``."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
needs_reference = pytest.mark.skipif(not os.path.isdir('/root/reference/kapre'), reason='/root/reference not present')


def _run(args):
    out = subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@needs_reference
def test_layer_constructors_and_get_config_match_the_reference():
    out = _run([os.path.join(HERE, 'golden', 'check_api_fidelity.py'), '150'])
    assert '0 disagreements' in out


@needs_reference
def test_oracle_agrees_with_the_reference_code_on_random_configurations():
    out = _run([os.path.join(HERE, 'golden', 'make_golden_ref.py'), '--fuzz', '120'])
    assert '0 disagreements' in out
