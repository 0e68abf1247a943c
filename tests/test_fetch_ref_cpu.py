"""oracle/fetch_ref.py: the recipe that puts the UNMODIFIED reference under the git-ignored oracle/_ref/ (so that it travels to
the GPU box like a built .so) and imports it under private module names."""
import os
import subprocess
import sys

import pytest

from oracle import fetch_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ref_dir_is_git_ignored_but_travels():
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in gi
    gpi = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gpi) or "oracle/_ref" not in open(gpi).read()
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    assert tracked == "", "reference sources must never be committed"


@pytest.mark.skipif(not os.path.isdir(fetch_ref.REF_SRC), reason="reference tree only exists in the build container")
def test_fetch_copies_verbatim():
    assert fetch_ref.fetch(verbose=False)
    for n in fetch_ref.FILES:
        assert open(os.path.join(fetch_ref.REF_SRC, n), "rb").read() == open(os.path.join(fetch_ref.REF_DST, n), "rb").read()
    assert fetch_ref.verify()


@pytest.mark.skipif(not fetch_ref.available(), reason="oracle/_ref not populated")
def test_import_under_private_names_and_runs():
    import torch
    before = {n: sys.modules.get(n) for n in ("module", "model", "utils", "dataset", "lipschitz")}
    ns = fetch_ref.import_ref()
    assert {n: sys.modules.get(n) for n in before} == before            # nothing shadows vae_song_b200's modules
    assert ns.model.LIDVAE.__module__ == "_ref_model" and ns.lipschitz.train_model.__module__ == "_ref_lipschitz"
    torch.manual_seed(0)
    m = ns.model.LIDVAE(dataset="pinwheel", icnn_channels=[16, 32], hidden_channels=[8, 4])
    ns.lipschitz.train_model(m, [(torch.randn(32, 2), None)] * 2, epochs=1, lr=1e-3, device="cpu")
    assert all(torch.isfinite(p).all() for p in m.parameters())
