"""CPU-side checks of front end B's host logic: the (q, f) tiling plan built inside the C-ABI library
against the oracle restatement (5 planes, 148 rows, 49 664 tiles), and state_dict
compatibility of the module mirrors with the reference classes.  No kernel is launched."""
import numpy as np
import torch

from oracle import qscan as OQ


def test_tiling_plan_matches_oracle_and_survey():
    from gw_whisper_b200.qfrontend import QScanB200
    q = QScanB200(1.0, 2048, [512, 512], [4, 128])
    assert (q.n_planes, q.n_rows, q.n_tiles) == (5, 148, 49664)
    pl = q.tiling_plan()
    ref = OQ.QScan(1.0, 2048, [512, 512], qrange=[4, 128]).tiling_plan()
    i = 0
    for ip, p in enumerate(ref):
        assert abs(p["q"] - pl["q"][ip]) < 1e-12
        for f, nt, ws in zip(p["freqs"], p["ntiles"], p["windowsize"]):
            assert (pl["freq"][i], pl["ntiles"][i], pl["windowsize"][i], pl["plane"][i]) == (f, nt, ws, ip)
            i += 1
    assert i == 148
    assert np.array_equal(pl["offset"], np.concatenate([[0], np.cumsum(pl["ntiles"])[:-1]]))
    rows_per_plane = np.bincount(pl["plane"]).tolist()
    assert rows_per_plane == [16, 25, 36, 43, 28]            # SURVEY.md Q1 tiling table


def test_unsupported_geometry_is_rejected():
    import pytest
    from gw_whisper_b200.qfrontend import QScanB200
    with pytest.raises(RuntimeError):
        QScanB200(2.0, 2048, [512, 512], [4, 128])            # 4096 samples: not this path
    with pytest.raises(RuntimeError):
        QScanB200(1.0, 2048, [512, 250], [4, 128])            # sides must be multiples of 64 in [64, 512]
    with pytest.raises(RuntimeError):
        QScanB200(1.0, 2048, [1024, 512], [4, 128])
    QScanB200(1.0, 2048, [128, 128], [4, 128])                # the MLGWSC-1/train.py geometry is supported


def test_adapter_state_dict_keys_match_reference_and_qtransform_buffers_are_ignored():
    from gw_whisper_b200.qfrontend import QTransformAdapter
    ours = QTransformAdapter()
    ref = OQ.QTransformAdapter()
    ref_keys = [k for k in ref.state_dict().keys() if not k.startswith("q_transform.")]
    assert list(ours.state_dict().keys()) == ref_keys
    sd = ref.state_dict()                                       # includes q_transform.* QTile buffers
    assert any(k.startswith("q_transform.") for k in sd)
    ours.load_state_dict(sd)                                    # strict load must accept them
    for k in ref_keys:
        assert torch.equal(ours.state_dict()[k], sd[k])


def test_classifier_mirror_structure_and_usr_mode():
    import pytest
    from gw_whisper_b200.qfrontend import GWWhisperClassifier
    with pytest.raises(TypeError):
        GWWhisperClassifier(torch.nn.Identity(), 2)            # needs the B200 encoder: no torch fallback
