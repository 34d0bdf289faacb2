"""LOBSTER month archives -> device-resident streams (replaces run_populate_database_from_zipped.py:50-109).

The reference walks a folder of ``_data_dwn_<..>__<TICKER>_<start>_<end>_<levels>.7z`` files in chronological order, extracts each
one with the external ``7z`` executable, runs ``populate_database`` on the trading days inside and deletes the CSVs again.  Same
steps here, with the Postgres insert replaced by the native packer (``pack_lobster`` -> ``DeviceDatabase.add_lobster_files``):

* the files are ordered by the part of the name behind ``__`` -- the dates, not the download number (:62-78);
* ticker, start date, end date and the number of levels come from the file name (:88-91);
* a ticker-day that is already in the database is not added again (:94-97 / populate_database.py:61-63);
* extraction uses the ``7z`` (or ``7za`` / ``7zr``) executable, as the reference does (:99); there is no 7z reader in Python's
  standard library, so a machine without the tool gets a loud error.  ``extractor`` is a hook for other containers;
* the extracted CSVs are removed after each archive (:106) -- only the files the extraction created, not every ``*.csv`` of the
  folder as the reference's ``delete_csvs`` does.
"""
from __future__ import annotations

import glob
import os
import re
import shutil
import subprocess
from datetime import datetime
from pathlib import Path
from typing import Callable, List, Optional, Tuple

ARCHIVE_GLOB = "*.7z"
_DAY_FILE = re.compile(r"^(?P<ticker>[^_]+)_(?P<date>\d{4}-\d{2}-\d{2})_\d+_\d+_message_(?P<levels>\d+)\.csv$")


def archive_order(fpaths: List[str]) -> List[str]:
    """Chronological order: sort by everything after the ``__`` of the name (run_populate_database_from_zipped.py:62-78)."""
    return sorted(fpaths, key=lambda f: f.split("__")[1])


def parse_archive_name(fpath) -> Tuple[str, datetime, datetime, int]:
    """``(ticker, start_date, end_date, n_levels)`` from ``..._<TICKER>_<YYYY-MM-DD>_<YYYY-MM-DD>_<levels>.7z`` (:88-91)."""
    parts = os.path.basename(str(fpath)).split("_")
    if len(parts) < 4:
        raise ValueError(f"not a LOBSTER archive name: {fpath}")
    try:
        return (parts[-4], datetime.strptime(parts[-3], "%Y-%m-%d"), datetime.strptime(parts[-2], "%Y-%m-%d"),
                int(parts[-1].split(".")[0]))
    except ValueError as e:
        raise ValueError(f"not a LOBSTER archive name: {fpath}") from e


def extract_7z(fpath, out_dir) -> None:
    """``7z x <archive> -o<dir>`` (:99).  Raises when no 7z executable is on PATH or the extraction fails."""
    exe = next((e for e in ("7z", "7za", "7zr") if shutil.which(e)), None)
    if exe is None:
        raise RuntimeError("no 7z / 7za / 7zr executable on PATH: install p7zip, or pass extractor= for another container")
    subprocess.run([exe, "x", str(fpath), "-o" + str(out_dir), "-y"], check=True, stdout=subprocess.DEVNULL)


def populate_from_archives(db, path_to_lobster_data, extractor: Optional[Callable] = None, keep_csvs: bool = False,
                           max_rows: Optional[int] = None, **pack_kw) -> List[int]:
    """Every ``*.7z`` month archive of the folder, oldest first, into ``db`` (a ``DeviceDatabase``); returns the stream ids added.

    One stream per ticker-day found in the archive whose date lies in the archive's [start, end] range (the archive holds
    exactly the trading days of its month, so no exchange calendar is needed to enumerate them).  ``pack_kw`` goes to
    ``pack_lobster`` (``step_us``, ``tie_order``, ...)."""
    extractor = extractor or extract_7z
    folder = Path(path_to_lobster_data)
    added: List[int] = []
    for fpath in archive_order(glob.glob(str(folder / ARCHIVE_GLOB))):
        ticker, start, end, n_levels = parse_archive_name(fpath)
        before = set(os.listdir(folder))
        extractor(fpath, folder)
        created = sorted(set(os.listdir(folder)) - before)
        try:
            days = []
            for name in sorted(os.listdir(folder)):
                m = _DAY_FILE.match(name)
                if not m or m["ticker"] != ticker or int(m["levels"]) != n_levels:
                    continue
                day = datetime.strptime(m["date"], "%Y-%m-%d")
                if start <= day <= end:
                    days.append((day, name))
            for day, name in days:
                if db.has_day(ticker, day):            # "already in database so not re-added"
                    continue
                books = glob.glob(str(folder / f"{ticker}_{day:%Y-%m-%d}_*_orderbook_{n_levels}.csv"))
                if not books:                           # get_book_and_message_paths, database_population_helpers.py:107-113
                    raise FileNotFoundError(f"Level {n_levels} data for ticker {ticker} on {day:%Y-%m-%d} not found in {folder}")
                added.append(db.add_lobster_files(ticker, day, folder / name, books[0], n_levels, max_rows=max_rows, **pack_kw))
        finally:
            if not keep_csvs:
                for name in created:
                    if name.endswith(".csv"):
                        os.remove(folder / name)
    return added
