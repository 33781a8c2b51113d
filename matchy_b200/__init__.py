"""matchy_b200 — B200-native implementation of matchy's log-scan hot path behind the reference's API.

    from matchy_b200 import Database, Extractor, processing
    db = Database.from_("threats.mxy").open()
    ex = Extractor.builder().extract_bitcoin(False).extract_ethereum(False).extract_monero(False).build()
    worker = processing.Worker.builder().extractor(ex).add_database("threats", db).build()
    for m in worker.process_bytes(open("access.log", "rb").read()): ...

Importing the package does not touch the GPU; creating a Database / Engine does, and fails loudly without one.
"""
from . import processing  # noqa: F401
from .builder import DatabaseBuilder, MatchMode  # noqa: F401
from .database import Database, DatabaseError, QueryResult  # noqa: F401
from .engine import ITEM_TYPE_NAMES, Engine, EngineError, RecordFormatter  # noqa: F401
from .extractor import Extractor, ExtractorError  # noqa: F401

__version__ = "0.1.0"
