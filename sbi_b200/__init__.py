"""sbi_b200: B200-native (sm_100a) density-estimator training and posterior evaluation,
drop-in at sbi's builder / estimator API.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
