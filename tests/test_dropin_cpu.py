"""CPU: the drop-in layer resolves the reference's import statements to this package (no GPU work is launched)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_statements_resolve_to_pdgn_b200(tmp_path):
    # a miniature of the reference tree: namespace packages utils/, models/ with one consumer module
    (tmp_path / "utils").mkdir()
    (tmp_path / "utils" / "misc.py").write_text("MARK = 'reference utils.misc'\n")
    (tmp_path / "models").mkdir()
    (tmp_path / "models" / "PDGNet_v2.py").write_text(textwrap.dedent("""
        from lib.pointops.functions import pointops
        from evaluation.evaluation_metrics import *
        from utils import chamfer_loss
        from utils import misc
        def get_edge_features(x, k, num=-1):
            return 'reference torch implementation'
        def get_edge_features_xyz(x, pc, k, num=-1):
            return 'reference torch implementation'
        class PDGNet_v2(object):
            def get_local_pair(self, pt1, pt2):
                return 'reference composition'
    """))
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from pdgn_b200 import dropin
        dropin.install(); dropin.install()
        import models.PDGNet_v2 as M
        import pdgn_b200.pointops, pdgn_b200.chamfer_loss, pdgn_b200.evaluation_metrics, pdgn_b200.edge_features
        assert M.pointops is pdgn_b200.pointops
        assert M.chamfer_loss is pdgn_b200.chamfer_loss
        assert M.compute_all_metrics is pdgn_b200.evaluation_metrics.compute_all_metrics
        assert M.misc.MARK == 'reference utils.misc'
        assert M.get_edge_features is pdgn_b200.edge_features.get_edge_features
        assert M.get_edge_features_xyz is pdgn_b200.edge_features.get_edge_features_xyz
        assert M.PDGNet_v2.get_local_pair.__name__ == '<lambda>'      # rebound to pdgn_b200.local_pair.get_local_pair
        import pointops_cuda
        for f in ['knnquery_cuda', 'grouping_forward_cuda', 'grouping_backward_cuda', 'nearestneighbor_cuda',
                  'interpolation_forward_cuda', 'interpolation_backward_cuda']:
            assert callable(getattr(pointops_cuda, f))
        print('dropin ok')
    """) % (ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "dropin ok" in r.stdout
