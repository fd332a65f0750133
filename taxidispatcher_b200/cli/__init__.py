"""Command-line twins of the reference's executables (file protocols of SURVEY.md section 8(b))."""
