class DataLoader:  # pragma: no cover - loader is out of scope
    pass
