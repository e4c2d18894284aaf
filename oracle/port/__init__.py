"""CPU restatement (oracle) of MultiVae's training-step algorithm.  TEST INFRASTRUCTURE ONLY."""
