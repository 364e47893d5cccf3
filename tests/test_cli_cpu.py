"""CPU: the host side of the reference-compatible CLI (cuda-recommender_b200/host/) without a GPU — usage text, the
reference's flag set (src/extras.cpp:46-141), the dataset loader of src/tools.cpp:3-85 on the on-disk format this repo
writes.  The compute path is not touched (no -CUDA)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda-recommender_b200", "host", "b200_recommender")


def _run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cuda-recommender_b200"), "all"])
    return CLI


def test_usage_lists_the_reference_flags(cli):
    out = _run([cli]).stdout
    assert "Usage:" in out
    for flag in ("-k rank", "-n threads", "-l lambda", "-t max_iter", "-T max_inner_iter", "-e epsilon", "-p do_predict",
                 "-q verbose", "-N do_nmf", "-CUDA"):
        assert flag in out, flag


def test_loader_reads_the_dataset_directory(cli, datagen, data_factory, tmp_path):
    d = data_factory("small")
    datagen.write_dataset(str(tmp_path), d)
    r = _run([cli, "-k", "7", "-l", "0.05", "-t", "2", "-T", "3", "-n", "2", str(tmp_path)])
    assert r.returncode == 0, r.stdout
    assert "[info] Loading R matrix" in r.stdout and "Picked Version: CCD!" in r.stdout
    assert "K = 7 | InnerIter = 3 | OuterIter = 2 | Threads = 2 | L = 0.050" in r.stdout
    assert "Total Time:" in r.stdout
    r = _run([cli, "-ALS", "-k", "3", str(tmp_path)])
    assert r.returncode == 0 and "Picked Version: ALS!" in r.stdout


def test_missing_dataset_fails_loudly(cli, tmp_path):
    r = _run([cli, "-k", "3", str(tmp_path / "nope")])
    assert r.returncode != 0


def test_usage_lists_the_round2_flags(cli):
    out = _run([cli]).stdout
    for text in ("-load", "-save", "-e switches the stop rule on"):
        assert text in out, text


def test_load_without_a_model_fails_loudly(cli, datagen, data_factory, tmp_path):
    """-load is the predict-only path (the reference's calculate_rmse_from_file): no model file -> an error, nothing computed."""
    datagen.write_dataset(str(tmp_path), data_factory("tiny"))
    r = _run([cli, "-load", str(tmp_path)])
    assert r.returncode != 0 and "can't open model file" in r.stdout
    assert not os.path.exists(os.path.join(str(tmp_path), "output"))


def test_load_rejects_a_truncated_model(cli, datagen, data_factory, tmp_path):
    datagen.write_dataset(str(tmp_path), data_factory("tiny"))
    with open(os.path.join(str(tmp_path), "model"), "wb") as f:
        f.write(b"\x05\x00\x00\x00")
    r = _run([cli, "-load", str(tmp_path)])
    assert r.returncode != 0
