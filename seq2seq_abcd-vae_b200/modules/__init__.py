# coding: utf-8
