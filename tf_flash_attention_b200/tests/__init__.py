"""Package-side verification and benchmark groups with the reference's command line
(`flash_attention/tests/`): `python -m tf_flash_attention_b200.tests.test_1d TestGroup.verify`."""
