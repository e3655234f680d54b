/* ncurses_stub.c -- TEST INFRASTRUCTURE.  The reference bundles samtools 0.1.18 (dependency/Linux/x64/samtools), which links
 * libncurses.so.5 / libtinfo.so.5 for its `tview` sub-command only.  Those libraries are not installed in this image, so the
 * binary would not even start.  This file provides the handful of curses symbols it imports as no-ops; built as
 * oracle/_ref/lib/libncurses.so.5 and libtinfo.so.5 (see oracle/Makefile, target `samtools`) it lets the UNMODIFIED reference
 * binary run `samtools faidx` here, which is what pins mir_prefer_b200/fastaindex.py (tests/golden/make_golden_faidx.py). */
void *stdscr = 0;
#define STUB(name) int name() { return 0; }
STUB(cbreak) STUB(delwin) STUB(endwin) STUB(init_pair) STUB(keypad) STUB(mvprintw) STUB(mvwprintw)
STUB(noecho) STUB(start_color) STUB(waddch) STUB(wattr_off) STUB(wattr_on) STUB(wborder) STUB(wclear) STUB(wgetch) STUB(wmove)
STUB(wrefresh)
void *initscr() { return 0; }
void *newwin() { return 0; }
