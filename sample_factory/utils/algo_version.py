"""sample_factory.utils.algo_version: the reference version this engine mirrors (utils/algo_version.py of sample-factory 2.1.3)"""
ALGO_VERSION = 83
