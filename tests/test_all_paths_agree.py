"""Every kernel path on one mixed, edge-heavy batch (tools/sanitize_paths.py): small units by
popcount and on the tensor cores, everything forced through the int8 Gram kernel, pipelines
of 2 / 5 / 40 groups, empty batches -- all must give the same records and tables."""
import os
import runpy

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_all_paths_agree_on_a_mixed_batch(gpu_ctx):
    runpy.run_path(os.path.join(ROOT, "tools", "sanitize_paths.py"), run_name="__main__")
