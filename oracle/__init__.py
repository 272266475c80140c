"""ORACLE — CPU restatements of the reference algorithm. Test infrastructure only."""
