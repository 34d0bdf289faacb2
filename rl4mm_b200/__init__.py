"""rl4mm_b200 -- a B200-native limit-order-book simulation step behind rl4mm's Exchange / OrderbookSimulator /
HistoricalOrderbookEnvironment interfaces.  See DESIGN.md."""

__version__ = "0.1.0"
