"""C1/C2: flag parity with reference run.py:7-20 and launcher compatibility."""
import pytest

from ddpt_b200.config import build_parser, parse_args


def test_reference_defaults():
    a = parse_args([])
    assert (a.exp_name, a.learning_rate, a.batch_size, a.epochs) == ("DDP_warmup", 0.1, 1024, 500)
    assert (a.warmup_epochs, a.warmup_type, a.seed_num) == (10, "linear", 42)
    assert (a.data_path, a.model_path) == ("./datasets/", "./experiment_model/")


def test_short_flags_and_equals_form():
    a = parse_args(["-n=exp", "-l", "0.5", "-b", "8", "-e", "3", "-w", "1", "-t", "cosine",
                    "-s", "7", "-d", "/d", "-p", "/p"])
    assert (a.exp_name, a.learning_rate, a.batch_size, a.epochs, a.warmup_epochs) == ("exp", 0.5, 8, 3, 1)
    assert (a.warmup_type, a.seed_num, a.data_path, a.model_path) == ("cosine", 7, "/d", "/p")


@pytest.mark.parametrize("spelling", ["--local_rank=3", "--local-rank=3"])
def test_both_launcher_spellings(spelling):
    assert parse_args([spelling]).local_rank == 3


def test_local_rank_from_env(monkeypatch):
    monkeypatch.setenv("LOCAL_RANK", "5")
    assert parse_args([]).local_rank == 5
    monkeypatch.delenv("LOCAL_RANK")
    assert parse_args([]).local_rank == 0


def test_bad_warmup_type_rejected():
    with pytest.raises(SystemExit):
        parse_args(["-t", "exponential"])


def test_every_reference_flag_present():
    opts = {s for a in build_parser()._actions for s in a.option_strings}
    for flag in ["-n", "--exp_name", "-l", "--learning_rate", "-b", "--batch_size", "-e", "--epochs",
                 "-w", "--warmup_epochs", "-t", "--warmup_type", "-s", "--seed_num", "-d", "--data_path",
                 "-p", "--model_path", "--local_rank"]:
        assert flag in opts
