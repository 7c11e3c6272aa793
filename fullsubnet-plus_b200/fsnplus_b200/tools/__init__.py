"""Command-line tools that mirror the reference's ``speech_enhance/tools``."""
